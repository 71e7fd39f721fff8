run() { # n extra-args env
n=$1; shift
env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 $EXTRA > gpurun_out/tmp_n$n.json 2> gpurun_out/tmp_n$n.err
python -c "
import json; d=json.load(open('gpurun_out/tmp_n$n.json')); print('n=$n', '$EXTRA', '$*', d['value'], d['ms_per_step'], d['e2e']['value'])"
}
EXTRA="" run 8 A=1
EXTRA="--overlap-allreduce" run 8 A=1
EXTRA="--overlap-allreduce" run 8 NCCL_MAX_CTAS=4
EXTRA="" run 8 NCCL_MAX_CTAS=8
EXTRA="--overlap-allreduce" run 4 NCCL_MAX_CTAS=4
EXTRA="" run 4 A=1
