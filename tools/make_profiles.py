#!/usr/bin/env python
"""Turn the outputs of tools/final_round.sh (gpurun_out/*_<tag>.*) into the tracked summaries under profiles/.

    python tools/make_profiles.py <tag> [<ncu-rep of the two-pass path>]      (FF_ROUND=r02 names the files)
"""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RND = os.environ.get("FF_ROUND", "r02")
tag = sys.argv[1]
rep = sys.argv[2] if len(sys.argv) > 2 else None
out = os.path.join(ROOT, "gpurun_out")
prof = os.path.join(ROOT, "profiles")


def short(name):
    return name.replace("void ", "").replace("ff::", "").split("(")[0]


rows = list(csv.reader(l for l in open(f"{out}/bench_launches_{tag}.csv") if l.startswith('"')))
hdr = rows[0]
ki, vi, gi, bi = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Grid Size", "Block Size"))
data = [(r[ki].replace("void ", "").replace("ff::", ""), float(r[vi].replace(",", "")) / 1000, r[gi], r[bi]) for r in rows[1:]]
start = max(i for i, d in enumerate(data) if d[0].startswith(("k_links_hist", "k_links_uniform")))   # the first kernel of a step
step = data[start:]
tot = sum(d[1] for d in step)
bench = json.loads(open(f"{out}/bench_{tag}.json").read().strip().splitlines()[-1])
frame = [i for i, d in enumerate(step) if "k_frame_merge" in d[0]]
if frame:
    first_sim = first_gather = frame[0]
    m0_names = "k_frame_merge, one launch"
else:
    first_gather = next(i for i, d in enumerate(step) if "k_merge_gather" in d[0])
    first_sim = next(i for i, d in enumerate(step) if "k_similarity" in d[0])
    m0_names = "k_similarity + k_keep_scan + k_merge_gather"
m0 = sum(d[1] for d in step[first_sim:first_gather + 1])
doc = [f"# {RND} — kernels of one bench step (C2), ncu launch list (final code of the round)", "",
       "Command (on the B200 box): `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline`",
       f"(raw list: `{RND}_bench_launches_raw.csv`; per-launch times under ncu are cold-cache and serialised — compare shares, not absolutes).", "",
       "| # | kernel | grid | block | time (us) | share |", "|---|---|---|---|---|---|"]
for i, d in enumerate(step):
    doc.append(f"| {i} | `{short(d[0])}` | {d[2]} | {d[3]} | {d[1]:.1f} | {100 * d[1] / tot:.1f} % |")
r = bench["roofline"]
doc += ["", f"Sum: {tot:.1f} us in {len(step)} launches.  Merge call #0 (the roofline kernel set, rows {first_sim}-{first_gather}: {m0_names}) "
        f"is {m0:.1f} us = {100 * m0 / tot:.1f} % of the step's kernel time; CUDA-event timing of the same launches inside bench.py, warm and "
        f"back to back: {r['kernel_us']:.1f} us -> {r['achieved']:.0f} GB/s algorithmic = {r['frac']:.3f} of the {r['peak']:.0f} GB/s peak ({r['peak_source']}).", "",
        f"Step wall time in the same bench run without ncu: {bench['ms_per_step'] * 1000:.0f} us for {len(step)} launches + 3 status reads; the difference to the "
        "kernel sum is host dispatch (Python + ctypes) that the GPU does not hide: the host reads the status block as soon as the deciding kernel "
        "has published it (ff_status_wait) and prepares the next call while the rest of that call runs (the frame-pipelined kernel sends it "
        "when the last frame is decided, a few microseconds before its end; the scan kernels before their gather).",
        "The second merge call closes merging: nothing crosses the threshold any more, its gather returns after reading the counter and the operator hands its inputs back.",
        "The last `k_merge_gather` compacts the pruned sequence (no averaging)."]
open(f"{prof}/{RND}_bench_step_launches.md", "w").write("\n".join(doc) + "\n")
subprocess.run(["cp", f"{out}/bench_launches_{tag}.csv", f"{prof}/{RND}_bench_launches_raw.csv"], check=True)
subprocess.run(["cp", f"{out}/bench_{tag}.json", f"{prof}/{RND}_bench_line.json"], check=True)
subprocess.run(["cp", f"{out}/bench_ref_{tag}.json", f"{prof}/{RND}_bench_reference_line.json"], check=True)
print("\n".join(doc[-6:]))

if rep:
    top = subprocess.run([sys.executable, f"{ROOT}/tools/ncu_top.py", rep, "8"], capture_output=True, text=True).stdout
    lines = subprocess.run([sys.executable, f"{ROOT}/tools/ncu_lines.py", rep, "14"], capture_output=True, text=True).stdout
    head = top.split("--- kernel 0")[0].rstrip()
    dram = []
    for blk in head.split("\n"):
        if "dram__bytes_read.sum" in blk or "dram__bytes_write.sum" in blk:
            dram.append(float(blk.split()[1]))
    total = sum(dram)
    doc = f"""# {RND} — ncu --set full, two-pass merge path, C2 (k_similarity, k_keep_scan, k_merge_gather) — final code of the round

Command: `ncu --set full --clock-control none --import-source on -k regex:"k_similarity|k_keep_scan|k_merge_gather" -s 6 -c 3 python tools/time_merge.py --cfg C2 --fused 0 --iters 2`
(per-launch times under ncu are cold-cache and serialised; the launch list in {RND}_bench_step_launches.md comes from bench.py itself)

```
{head}
```

DRAM traffic of the call (reads + writes of the three kernels): **{total:.1f} MB** for 455.0 MB algorithmic — the second read of
`hidden_states` is the difference; `traffic.json` carries the per-launch figure bench.py reports.

## Stall samples per source line (the three kernels together)

```
{lines.rstrip()}
```

Every hot line waits on `long_scoreboard` (global loads): the path is latency bound, not issue bound.  The three kernels
are the ones of r01 (DESIGN.md section 5 has what moved them and what did not); this round's work on the merge stage went
into the read-once kernel ({RND}_read_once_kernel.md), which ends level with this path.
"""
    open(f"{prof}/{RND}_two_pass_ncu_full.md", "w").write(doc)
    t = json.load(open(f"{prof}/traffic.json"))
    t["C2"]["dram_bytes_per_launch"] = int(round(total * 1e6, -5))
    t["C2"]["kernel"] = "ff_merge_layer call #0, two-pass path: k_similarity + k_keep_scan + k_merge_gather (ncu --set full, " + RND + "_two_pass_ncu_full.md; writes still in L2 at kernel end are not counted)"
    json.dump(t, open(f"{prof}/traffic.json", "w"), indent=1)
    print("traffic", total)
