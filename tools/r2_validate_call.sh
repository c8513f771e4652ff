#!/bin/bash
# One-GPU validation call: the whole `-m gpu` suite, smoke(), and a short bench line with the GPU reference leg.
#     gpurun --timeout 1500 -- 'bash tools/r2_validate_call.sh [tag]'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-validate}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 700 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.err
python - "$TAG" <<'PY'
import json, sys
tag = sys.argv[1]
d = json.loads(open(f"gpurun_out/{tag}_bench.json").read().strip().splitlines()[-1])
print("steps/s", d["value"], "ms/step", d["ms_per_step"], "attn ms", d["roofline"]["avg_launch_ms"], "frac", d["roofline"]["frac"])
print("pipeline", d["pipeline"] and (d["pipeline"]["seconds"], d["pipeline"]["frames_per_sec"]), "clocks", d["clocks"])
print("sha", d.get("latents_sha256"))
for k in d["top_kernels"]:
    print("  ", k)
g = d.get("gpu_reference") or {}
print("gpu_reference", {k: g.get(k) for k in ("block_ms", "self_attention_fa2_ms", "self_attention_sdpa_ms", "linear_cxc_ms", "ffn_ms", "speedup", "unavailable")})
PY
