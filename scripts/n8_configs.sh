set -x
run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus 8 "$@" > gpurun_out/${tag}.json 2> gpurun_out/${tag}.err; tail -2 gpurun_out/${tag}.err | cut -c1-300; }
run c4_cfg3_n8 --genomes 3 --divergence 1.3 --steps 3 --warmup 3
run c4_cfg4_n8 --genomes 5 --divergence 12 --steps 2 --warmup 3
run c4_cfg5_n8 --steps 3 --warmup 3
python - <<'PY'
import json
for f in ("c4_cfg3_n8","c4_cfg4_n8","c4_cfg5_n8"):
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],1), round(d["value"]/1e9,2), round(d["e2e"]["ms_per_step"],1), d["config"]["blocks"], d["config"]["blocks_sha1"][:10], d["config"].get("phase_ms_rank0"), d["roofline"]["kernel_ms_per_step"])
        print("  ", d.get("graph_stage_phase_ms"))
    except Exception as e:
        print(f, "failed", e)
PY
