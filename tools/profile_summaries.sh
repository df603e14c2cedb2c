#!/bin/bash
# usage: tools/profile_summaries.sh <tag in gpurun_out, e.g. r03z> <name under profiles, e.g. r02>
# Turns gpurun_out/<tag>_launches.csv and gpurun_out/<tag>_prof.ncu-rep into the committed summaries under profiles/.
set -e
TAG=$1; OUT=$2; ROOT=$(cd "$(dirname "$0")/.." && pwd); TMP=$(mktemp -d)
python - "$ROOT/gpurun_out/${TAG}_launches.csv" "$ROOT/profiles/${OUT}_launches.txt" <<'PY'
import csv, collections, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith('==')))
h = rows[0]; ik = h.index('Kernel Name'); iv = h.index('Metric Value'); iu = h.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= iv: continue
    v = float(r[iv].replace(',', '')); u = r[iu]
    v = v / 1e6 if u == 'ns' else v / 1e3 if u == 'us' else v * 1e3 if u == 's' else v
    agg[r[ik]][0] += 1; agg[r[ik]][1] += v
tot = sum(v[1] for v in agg.values())
out = ["# ncu --metrics gpu__time_duration.sum --clock-control none -c 40: python bench.py --steps 2 --warmup 3 --no-cpu --no-other", "# kernel, launches, total ms, share"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]): out.append("%-90s %3d %12.3f ms %7.3f %%" % (k[:90], v[0], v[1], 100 * v[1] / tot))
open(sys.argv[2], 'w').write("\n".join(out) + "\n")
PY
ncu -i "$ROOT/gpurun_out/${TAG}_prof.ncu-rep" --page details > "$ROOT/profiles/${OUT}_ncu_full_bl_run_kernel.txt"
ncu -i "$ROOT/gpurun_out/${TAG}_prof.ncu-rep" --page source --csv > "$TMP/src.csv"
(cd "$TMP" && cuobjdump -xelf all "$ROOT/bayeslands_b200/libbayeslands_b200.so" > /dev/null && nvdisasm --print-line-info bl_v512.sm_100a.cubin > v512.dis)
(cd "$ROOT/bayeslands_b200/csrc" && python "$ROOT/tools/ncu_stalls.py" "$TMP/src.csv" "$TMP/v512.dis" bl_run_kernelILi8 60 > "$ROOT/profiles/${OUT}_stalls_by_line.txt")
(echo "# cuobjdump -sass of bl_v512.sm_100a.cubin (libbayeslands_b200.so): TMA bulk copies (UBLKCP), mbarrier operations (SYNCS) and proxy fences of the compact layout";
 cuobjdump -sass "$TMP/bl_v512.sm_100a.cubin" | grep "UBLKCP\|SYNCS\|FENCE.VIEW.ASYNC" | sed 's/^ *\/\*[0-9a-f]*\*\/ *//; s/ *\/\*.*//' | sort | uniq -c | sort -k1 -n -r | head -20) > "$ROOT/profiles/${OUT}_sass_tma.txt"
ncu -i "$ROOT/gpurun_out/${TAG}_prof.ncu-rep" --page raw --csv | python -c "
import csv, sys, json
rows = list(csv.reader(sys.stdin)); h = rows[0]; v = rows[2]
g = lambda k: float(v[h.index(k)])
rd = g('dram__bytes_read.sum') * 1e9; wr = g('dram__bytes_write.sum') * 1e9
d = dict(source='profiles/${OUT}_ncu_full_bl_run_kernel.txt (ncu --set full --clock-control none, python bench.py --batch 296 --steps 1 --warmup 3 --no-cpu --no-other)', capture='${OUT} (gpurun ${TAG})', N=12727, B=296, mean_steps=240.118, dram_bytes_read=rd, dram_bytes_write=wr, kernel_ms=g('gpu__time_duration.sum'))
d['dram_bytes_per_node_step'] = (rd + wr) / (d['B'] * d['mean_steps'] * d['N'])
json.dump(d, open('$ROOT/profiles/ncu_traffic.json', 'w'), indent=1); print(d['kernel_ms'], d['dram_bytes_per_node_step'], rd / 1e9, wr / 1e9)
"
rm -rf "$TMP"
