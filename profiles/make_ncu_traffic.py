"""profiles/ncu_traffic.json (read by bench.py for `roofline.traffic`) from an ncu CSV of per-launch DRAM bytes.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        -k regex:conv3x3x3_tc --launch-skip 19 --launch-count 19 --log-file gpurun_out/r02_conv_dram_b25.csv \
        python profiles/run_ncu_forward.py 25
    python profiles/make_ncu_traffic.py gpurun_out/r02_conv_dram_b25.csv conv3x3x3_tc <algorithmic bytes per launch> "<source note>"

The capture is one SwinUNETR forward of the batch bench.py uses (25 windows of 96^3): the 19 conv3x3x3_tc launches of a forward,
averaged -- the same population as bench.py's `roofline.launches` average.
"""
import csv
import json
import os
import sys


def main():
    path, kernel, algo, note = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if "Metric Name" in r)
    i_id, i_name, i_metric, i_unit, i_val = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = {}
    for r in rows:
        if r is hdr or kernel not in r[i_name] or not r[i_metric].startswith("dram__bytes"):
            continue
        per.setdefault(r[i_id], 0.0)
        per[r[i_id]] += float(r[i_val].replace(",", "")) * scale[r[i_unit]]
    if not per:
        raise SystemExit(f"no {kernel} launches in {path}")
    out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json")
    data = json.load(open(out_path)) if os.path.exists(out_path) else {}
    data[kernel] = {"dram_bytes_per_launch": sum(per.values()) / len(per), "algorithmic_bytes_per_launch": algo, "launches": len(per),
                    "source": note}
    json.dump(data, open(out_path, "w"), indent=1)
    print(json.dumps(data[kernel]))


if __name__ == "__main__":
    main()
