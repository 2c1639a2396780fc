# All-reduce of the gradient arena under different NCCL settings (tools/nccl_probe.py).  Keep the per-setting timeout SHORT:
# a setting NCCL refuses (NVLS on 2 ranks) or a rank hanging in teardown otherwise burns GPU-minutes on every GPU of the call.
mkdir -p gpurun_out
N=${1:-2}
run() { tag="$1"; shift; env "$@" timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/nccl_probe.py "$tag" 2>&1 | grep -E "NCCLPROBE|Error|error" | head -3; }
run default NCCL_DEBUG=WARN
run minch16 NCCL_MIN_NCHANNELS=16
run minch32 NCCL_MIN_NCHANNELS=32
run maxch4 NCCL_MAX_NCHANNELS=4
run ring NCCL_ALGO=Ring
run tree NCCL_ALGO=Tree
run nvls NCCL_ALGO=NVLS
run ll128 NCCL_PROTO=LL128
run simple NCCL_PROTO=Simple
run ring_minch32 NCCL_ALGO=Ring NCCL_MIN_NCHANNELS=32
