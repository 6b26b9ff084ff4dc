# builder tool: 8-GPU torchrun of bench.py (detect leg with 2 lanes and with 1 lane)
for lanes in 2 1; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$lanes bench.py --gpus 8 --steps 6 --warmup 3 --detect-lanes $lanes > gpurun_out/b8_$lanes.json 2> gpurun_out/b8_$lanes.err; echo rc=$?
python - <<PY
import json
for l in open("gpurun_out/b8_$lanes.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["n_gpus"], round(d["value"]/1e6,1), round(d["e2e"]["value"]/1e6,1), round(d["detect"]["value"],1), round(d["detect"]["ms_per_batch"],2), d["detect"]["detections_total"], d["detect"]["config"]["lanes"])
PY
done
nproc; free -g | head -2
