"""tests/golden/bench_candidates.json from a SACB_TRACE_DDS trace of the bench (frame records of 882000 samples): every 40th
evaluated candidate of ONE frame's search, in order -- the candidates bench.py's cpu_baseline / reference arm time the CPU on.
usage: SACB_TRACE_DDS=gpurun_out/trace.jsonl python bench.py ... ; python tools/make_bench_candidates.py gpurun_out/trace.jsonl"""
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1])]
rows = [r for r in rows if r["frame_n"] == 882000]   # the timed frames (the warm-up runs 2-s excerpts)
first = []
for r in rows:                                   # the first frame written (steps restart at 1 for the next one)
    if first and r["step"] <= first[-1]["step"]:
        break
    first.append(r)
step = max(1, len(first) // 25)
sel = first[step // 2::step][:25]
out = {"_comment": "every %d-th of the %d candidates the GPU arm evaluated for one 20-s frame of the bench stream (seed 3), --best --opt-cfg=dds,128; "
                   "x = the 56 searched coefficients (profile indices 0..55)" % (step, len(first)),
       "steps": [r["step"] for r in sel], "cost": [r["cost"] for r in sel], "x": [r["x"] for r in sel]}
json.dump(out, open("tests/golden/bench_candidates.json", "w"))
print(len(first), "candidates ->", len(sel))
