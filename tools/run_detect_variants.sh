# builder tool: detection leg of bench.py under lane / lazy-compaction settings (one summary line each)
summ() { python -c "
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l)['detect']; print(sys.argv[2], round(d['value'],1), round(d['ms_per_batch'],2), d['detections_total'], d['host_syncs_per_batch'], d['stage_ms'])
" $1 "$2"; }
python -m pytest tests/test_gpu_cascade.py -m gpu -x -q -k "contrast or lazy" 2>&1 | tail -2
for lanes in 1 2 3 4; do
  python bench.py --steps 12 --warmup 3 --no-cpu-baseline --detect-lanes $lanes > gpurun_out/dv.json 2> gpurun_out/dv.err || tail -5 gpurun_out/dv.err
  summ gpurun_out/dv.json "lanes=$lanes"
done
for thr in 2048 8192; do
  HGSFA_LAZY_THRESHOLD=$thr python bench.py --steps 12 --warmup 3 --no-cpu-baseline --detect-lanes 2 > gpurun_out/dv.json 2> gpurun_out/dv.err || tail -5 gpurun_out/dv.err
  summ gpurun_out/dv.json "lanes=2 lazy=$thr"
done
