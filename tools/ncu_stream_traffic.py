"""profiles/r2_stream_traffic.json from an `ncu --set full` capture of the stream kernel launched by bench.py itself:
    ncu --set full --clock-control none -k regex:k_table_add_sample -c 1 -o gpurun_out/r2_stream_bench -f python bench.py --timed-only --steps 1 --warmup 1
    python tools/ncu_stream_traffic.py gpurun_out/r2_stream_bench.ncu-rep <records> <table_keys>
bench.py fills roofline.traffic from this file when its launch has the same shape."""
import csv
import json
import subprocess
import sys
from pathlib import Path

rep, records, keys = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
d = dict(zip(hdr, rows[2]))


def val(name):
    v = float(d[name].replace(",", ""))
    u = units[hdr.index(name)].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


out = {"kernel": d["Kernel Name"][:60], "records": records, "table_keys": keys, "dram_bytes": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")),
       "dram_bytes_read": int(val("dram__bytes_read.sum")), "dram_bytes_write": int(val("dram__bytes_write.sum")), "gpu_time_ms_under_ncu": float(d["gpu__time_duration.sum"].replace(",", "")) *
       {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(units[hdr.index("gpu__time_duration.sum")].lower().replace("second", "s").replace("nsecond", "ns"), 1),
       "source": "ncu --set full --clock-control none of the launch inside `bench.py --timed-only --steps 1 --warmup 1` (first k_table_add_sample launch)"}
Path(__file__).resolve().parent.parent.joinpath("profiles", "r2_stream_traffic.json").write_text(json.dumps(out, indent=1) + "\n")
print(out)
