"""profiles/rNN_traffic.json from the ncu summaries (summarize_ncu.py output): DRAM bytes per launch and pipe
utilisation of the dominant kernel of every config; bench.py copies them into roofline.traffic / roofline.ncu_*.
    python profiles/make_traffic.py r02 c2=profiles/r02_ncu_full_c2.txt[:launch_divisor] ... > profiles/r02_traffic.json
launch_divisor: iterations inside one captured launch (the fused ensemble run holds several)."""
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for spec in sys.argv[2:]:
    name, path = spec.split("=")
    div = 1.0
    if ":" in path:
        path, d = path.split(":")
        div = float(d)
    txt = open(path).read()

    def metric(key):
        m = re.search(r"^\s+" + re.escape(key) + r"\s+([0-9.]+)\s*(\S*)", txt, re.M)
        return (float(m.group(1)), m.group(2)) if m else (None, "")

    rd, ru = metric("dram__bytes_read.sum")
    wr, wu = metric("dram__bytes_write.sum")
    out[name] = {
        "kernel": re.search(r"^kernel: (.*?) \|", txt, re.M).group(1)[:90],
        "dram_bytes_per_launch": (rd * UNIT[ru] + wr * UNIT[wu]) / div,
        "iterations_in_captured_launch": div,
        "source": f"ncu --set full --clock-control none, {path}",
        "tensor_pipe_active_pct": metric("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")[0],
        "fma_pipe_active_pct": metric("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed")[0],
        "issue_active_pct": metric("sm__issue_active.avg.pct_of_peak_sustained_elapsed")[0],
    }
print(json.dumps(out, indent=1))
