timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_r1_n8.json 2> gpurun_out/bench_r1_n8.err
echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_n8.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['launch'])"
