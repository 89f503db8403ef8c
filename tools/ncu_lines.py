"""Per CUDA source line: stall samples (cuda,sass view of an .ncu-rep), top lines + stall reason split."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; hdr = None
agg = collections.Counter(); text = {}; reasons = collections.defaultdict(collections.Counter); instr = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name" or hdr is None: continue
    if len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    ln = r[0]
    if not ln.isdigit(): continue
    key = (cur_file, int(ln))
    # rows with an Address are SASS rows under the line; the line row itself carries aggregated numbers
    if r[2] in ("", "-"):
        try: agg[key] += int(r[4])
        except ValueError: pass
        text[key] = r[1].strip()[:110]
        for k in hdr:
            if k.startswith("stall_") and "Not Issued" not in k:
                try: reasons[key][k] += int(d[k])
                except (ValueError, KeyError): pass
        try: instr[key] += int(d["Instructions Executed"])
        except (ValueError, KeyError): pass
tot = sum(agg.values()) or 1
print("total samples", tot, " total warp instr", sum(instr.values()))
for key, v in agg.most_common(topn):
    rs = ", ".join(f"{k[6:]}={n}" for k, n in reasons[key].most_common(3))
    print(f"{v:7d} {100*v/tot:5.1f}%  inst={instr[key]:9d} {key[0]}:{key[1]:4d}  {text.get(key,'')}\n{'':16s}[{rs}]")
