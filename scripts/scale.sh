for n in 4 8; do
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_r1_n$n.json 2> gpurun_out/bench_r1_n$n.err
echo "n=$n rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_n$n.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['launch'], d['roofline']['frac'])"
done
