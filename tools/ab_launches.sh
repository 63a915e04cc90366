#!/bin/bash
# A/B of environment switches on ONE box: per-launch device times of a finest-scale training step under each variant.
#   bash tools/ab_launches.sh "<tag>:<ENV=1 ENV2=0>" ...     (tag "base" with no env = defaults)
mkdir -p gpurun_out
for spec in "$@"; do
  tag=${spec%%:*}; envs=${spec#*:}
  [ "$envs" == "$spec" ] && envs=""
  env $envs ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/ab_${tag}.csv python tools/profile_step.py --scale ${AB_SCALE:-4} --steps 2 > gpurun_out/ab_${tag}.log 2>&1
done
python - "$@" <<'PY'
import csv, re, sys
def load(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    names = [r['Kernel Name'] for r in rows]
    marks = [i for i, n in enumerate(names) if 'qsample_mix' in n]
    seg = rows[marks[-1]:]
    return [(re.sub(r"\(.*", "", r['Kernel Name']).replace("sinddm::", "").replace("<unnamed>::", "").replace("void ", "")[:28],
             float(r['Metric Value'].replace(',', '')) / 1e3) for r in seg]
tags = [s.split(':')[0] for s in sys.argv[1:]]
data = {t: load(f'gpurun_out/ab_{t}.csv') for t in tags}
print("total ms:", {t: round(sum(v for _, v in d) / 1e3, 3) for t, d in data.items()})
for kern in ("tc_conv", "tc_wgrad", "dw5x5", "colsum", "final_conv", "simt"):
    print(kern, {t: round(sum(v for n, v in d if kern in n) / 1e3, 3) for t, d in data.items()})
print("tc_conv launches (us):")
cols = [[v for n, v in data[t] if 'tc_conv' in n] for t in tags]
print("   " + "".join(f"{t:>10s}" for t in tags))
for i in range(max(len(c) for c in cols)):
    print(f"{i:2d} " + "".join(f"{c[i]:10.1f}" if i < len(c) else " " * 10 for c in cols))
PY
