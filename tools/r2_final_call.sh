#!/bin/bash
# Final one-GPU validation of a round: whole `-m gpu` suite, smoke(), a bench line, the 81-frame VAE, two kernel timings.
#     gpurun --timeout 1800 -- 'bash tools/r2_final_call.sh [tag]'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-final}
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 400 gpurun_out/${TAG}_bench.err
python - "$TAG" <<'PY'
import json, sys
tag = sys.argv[1]
d = json.loads(open(f"gpurun_out/{tag}_bench.json").read().strip().splitlines()[-1])
print("steps/s", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "attn ms", d["roofline"]["avg_launch_ms"], "frac", d["roofline"]["frac"])
print("pipeline", d["pipeline"]["seconds"], d["pipeline"]["frames_per_sec"], "clocks", d["clocks"])
print("sha", d.get("latents_sha256"), "cpu_baseline", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("cores"))
for k in d["top_kernels"]:
    print("  ", k)
g = d.get("gpu_reference") or {}
print("gpu_reference", {k: g.get(k) for k in ("block_ms", "self_attention_fa2_ms", "self_attention_sdpa_ms", "linear_cxc_ms", "ffn_ms", "speedup", "unavailable")})
PY
timeout 300 python tools/vae_bench.py --frames 81 > gpurun_out/${TAG}_vae_81f.json 2>&1
python - "$TAG" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/{sys.argv[1]}_vae_81f.json").read().strip().splitlines()[-1])
print("VAE 81f: enc", d["enc"]["ms"], d["enc"]["conv_tflops"], "dec", d["dec"]["ms"], d["dec"]["conv_tflops"])
print(d["dec"]["top"])
PY
for pair in 0 1; do VCOF_GEMM_2CTA=$pair tests/native/kbench gemm 75600 5120 5120 2 10; done
python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
from videocof_b200 import ops
qkv = torch.randn(3, 14400, 1152, device="cuda").bfloat16()
ops.vae_attn(qkv, 384); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ops.vae_attn(qkv, 384)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5 / 3
print("vae_attn ms per 720p latent frame", ms, "TF/s", 4 * 14400 ** 2 * 384 / ms / 1e9)
PY
