#!/usr/bin/env python
"""Dev tool: turn the ncu CSVs of an evidence run into profiles/ncu_flops.json and profiles/ncu_traffic.json, stamped with
the sha256 of the library they were captured from (so file + SASS of the step kernels).  bench.py reports the executed
flops / DRAM traffic only when the library it loads carries the same hash.

    python tools/ncu_summarize.py TAG LIB KEY=flops.csv[,full.csv] ...
    e.g. python tools/ncu_summarize.py r2f gym-solarpvder-environment_b200/csrc/libpvder_b200.so \\
             model_1=profiles/r2f_flops_1ph.csv,profiles/r2f_step_kernel_1ph_ncu_full.csv

flops.csv: `ncu --metrics smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul,fp64}_pred_on.sum ... --csv` of ONE launch of
1,048,576 envs x 30 sub-steps;  full.csv: `ncu -i x.ncu-rep --page raw --csv` of a --set full capture of one launch, or its tools/ncu_compact.py form."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import library_hashes  # noqa: E402

ENVS, SUB = 1 << 20, 30


def metric_rows(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    if "Metric Name" in hdr:                       # long format (--metrics ... --csv)
        i, j = hdr.index("Metric Name"), hdr.index("Metric Value")
        return {r[i]: float(r[j].replace(",", "")) for r in rows[1:] if len(r) > j}
    if hdr[:3] == ["metric", "unit", "value"]:     # compact format (tools/ncu_compact.py)
        out, units = {}, {}
        for k, u, v in rows[1:]:
            units[k] = u
            try:
                out[k] = float(v.replace(",", ""))
            except ValueError:
                pass
        return out, units
    vals = rows[2]                                 # wide format (--page raw --csv): header, units, values
    out = {}
    for k, v in zip(hdr, vals):
        try:
            out[k] = float(v.replace(",", ""))
        except ValueError:
            pass
    return out, dict(zip(hdr, rows[1]))


def main():
    tag, lib = sys.argv[1], sys.argv[2]
    hashes = library_hashes(lib)
    flops = {"_doc": "FP64 flops ONE half-cycle sub-step of one env executes in the step kernels of the library with the hashes "
                     "below, from ncu instruction counters (2 x smsp__sass_thread_inst_executed_op_dfma_pred_on + ..dadd.. + "
                     "..dmul.., one launch of 1,048,576 envs x 30 sub-steps); bench.py reports it as roofline.achieved/frac",
             "tag": tag, "n_sim": 15, **hashes}
    traffic = {"_doc": "dram__bytes_read.sum + dram__bytes_write.sum of ONE step_kernel launch at 1,048,576 envs, from the ncu "
                       "--set full captures named in 'source' (bench.py scales by envs/2^20: roofline.traffic)",
               "tag": tag, "envs": ENVS, **hashes}
    for spec in sys.argv[3:]:
        key, _, files = spec.partition("=")
        files = files.split(",")
        m = metric_rows(files[0])
        dfma = m["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"]
        dadd = m["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
        dmul = m["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"]
        f64 = m.get("smsp__sass_thread_inst_executed_op_fp64_pred_on.sum", dfma + dadd + dmul)
        flops[key] = {"flop_per_sub_step": round((2 * dfma + dadd + dmul) / (ENVS * SUB), 1),
                      "fp64_inst_per_sub_step": round(f64 / (ENVS * SUB), 1), "source": os.path.relpath(files[0], ROOT)}
        if len(files) > 1:
            vals, units = metric_rows(files[1])
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            b = sum(vals[k] * scale[units[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            traffic[key] = {"bytes": int(b), "source": os.path.relpath(files[1], ROOT)}
    json.dump(flops, open(os.path.join(ROOT, "profiles", "ncu_flops.json"), "w"), indent=1)
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(flops, indent=1))
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
