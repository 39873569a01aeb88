#!/usr/bin/env bash
# GPU session r02n: per-launch durations of a wavefront at 1 part and at part 0 of 8 (ncu launch lists; shares, not absolutes) + queue sizes per bounce
set -u
O=gpurun_out; mkdir -p $O
for np in 1 8; do
  timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:k_intersect_staged -s 9 -c 9 --csv --log-file $O/r02n_launches_c4_parts$np.csv python scripts/profile_target.py c4 2 8 $np TravChunk=32 > $O/r02n_ncu_$np.log 2>&1; echo "parts $np rc=$?"
done
python - <<'PY'
import csv
for np_ in (1, 8):
    rows = list(csv.reader(open(f"gpurun_out/r02n_launches_c4_parts{np_}.csv")))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r); hd = rows[h]
    kn, mn, mv, mu, iid = hd.index("Kernel Name"), hd.index("Metric Name"), hd.index("Metric Value"), hd.index("Metric Unit"), hd.index("ID")
    per = {}
    for r in rows[h + 1:]:
        if len(r) > mv: per.setdefault(int(r[iid]), {})[r[mn]] = (float(r[mv].replace(",", "")), r[mu])
    for i in sorted(per):
        d = per[i]; t, u = d["gpu__time_duration.sum"]; t *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(u, 1e-6)
        print(np_, "launch", i, "ms %.3f" % t, "warp inst %.3e" % d["smsp__inst_executed.sum"][0], "lsu wavefronts %.3e" % d["l1tex__data_pipe_lsu_wavefronts.sum"][0], "lanes/inst %.2f" % d["smsp__thread_inst_executed_per_inst_executed.ratio"][0])
PY
