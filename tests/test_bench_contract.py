"""The reference arm of bench.py runs on the host cores: its JSON line must carry the keys of the measurement contract
(the GPU arm prints the same keys plus roofline / clocks; it is exercised on the GPU box)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "vision_tokens/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("vision tokens/sec through merge+prune")
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["n_gpus"] == 1 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("C2:") and d["data"] == "synthetic" and d["dtype"] == "bf16"
    # "reference" = the unmodified framefusion/main.py was found ($FF_REFERENCE_DIR, /root/reference or baseline/_ref),
    # "port" = the torch-CPU restatement of it stood in
    from oracle import ref_locate
    want = "reference" if ref_locate.find_reference_dir() else "port"
    assert d["cpu_baseline"]["kind"] == want and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert set(d["config"]) == {"workload", "seq_len", "calls_per_step", "l2"}
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
