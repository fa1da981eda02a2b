#!/bin/bash
# A/B timing of variant builds on the GPU box: tools/ab.sh OUT "v0 v1 ..." [quick_bench args via QB_* env]
# (variant "main" = the in-tree libpms.so).  NCU=1 adds a per-kernel duration list (ncu, gpu__time_duration only).
out=$1; shift
vars=$1; shift
size=${AB_SIZE:-256}; kinds=${AB_KINDS:-sgrid,hex8}
for v in $vars; do
  lib=percept_b200/lib/variants/libpms_$v.so; [ "$v" = main ] && lib=percept_b200/lib/libpms.so
  echo "=== $v" >> $out
  PMS_LIB=$PWD/$lib python tools/quick_bench.py $size $kinds >> $out 2>&1
  if [ -n "$NCU" ]; then
    PMS_LIB=$PWD/$lib ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"${AB_KREGEX:-k_elem_pipe|k_sgrid_tile|k_grad_gather|k_metric_multi|k_sgrid_seam}" --csv \
      --log-file gpurun_out/ab_ncu_$v.csv python tools/quick_bench.py $size $kinds > /dev/null 2>&1
    python - "$v" >> $out <<'PY'
import csv, sys, collections, re
v = sys.argv[1]
rows = [r for r in csv.reader(open(f"gpurun_out/ab_ncu_{v}.csv")) if len(r) > 10]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
acc = collections.defaultdict(list)
for r in rows[1:]:
    try: acc[re.sub(r"\(.*", "", r[ix["Kernel Name"]])].append(float(r[ix["Metric Value"]].replace(",", "")) / (1e3 if r[ix["Metric Unit"]] in ("ns", "nsecond") else 1))
    except Exception: pass
for k, t in acc.items():
    t = sorted(t); print(f"   ncu {k[:60]:60s} n={len(t):3d} median {t[len(t)//2]:9.1f} us  min {t[0]:9.1f}")
PY
  fi
done
cat $out
