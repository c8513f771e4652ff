cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 700 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c14_bench.json 2> gpurun_out/r2c14_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c14_bench.json").read().strip().splitlines()[-1])
print("steps/s", d["value"], "ms/step", d["ms_per_step"], "attn ms", d["roofline"]["avg_launch_ms"], "frac", d["roofline"]["frac"])
print("pipeline", d["pipeline"]["seconds"], d["pipeline"]["frames_per_sec"], d["pipeline"]["edit_frames_sha256"])
print("sha", d.get("latents_sha256"))
for k in d["top_kernels"]: print("  ", k)
print(d["gpu_reference"]["speedup"])
PY
timeout 300 python tools/vae_bench.py --frames 81 > gpurun_out/r2c14_vae_81f.json 2>&1; tail -c 900 gpurun_out/r2c14_vae_81f.json; echo
for i in 1 2 3; do tests/native/kbench attn 75600 75600 40 3; VCOF_ATTN_EMU=3 tests/native/kbench attn 75600 75600 40 3; done
