# N-GPU lease: host-ingest probe, then the default bench line (config 2, pose shards) at N and at 1 for the e2e scaling ratio.
# Usage: gpurun --gpus N -- 'bash tools/gpu_e2e_scaling.sh N TAG'
cd /root/repo; N=${1:-8}; TAG=${2:-r2e}
mkdir -p gpurun_out
lscpu | grep -i "numa\|model name\|^CPU(s)" > gpurun_out/${TAG}_host.txt 2>&1
nvidia-smi topo -m >> gpurun_out/${TAG}_host.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/d2h_probe.py 2>/dev/null | grep "^rank" | sort > gpurun_out/${TAG}_d2h_n$N.txt
cat gpurun_out/${TAG}_d2h_n$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 200 --cpu-frames 0 > gpurun_out/${TAG}_bench_cfg2_n$N.json 2> gpurun_out/${TAG}_bench_cfg2_n$N.err
python bench.py --steps 200 --cpu-frames 0 > gpurun_out/${TAG}_bench_cfg2_n1.json 2> gpurun_out/${TAG}_bench_cfg2_n1.err
for f in gpurun_out/${TAG}_bench_cfg2_n$N.json gpurun_out/${TAG}_bench_cfg2_n1.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=d['e2e']
    print(sys.argv[1].split('/')[-1], 'value %.0f e2e %.0f call16 %.0f pageable %.0f ms/step %.3f' % (d['value'], e['value'], e.get('call16_value',0), e.get('pageable_value',0), d['ms_per_step']), d.get('azimuth_sharded'))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
done
