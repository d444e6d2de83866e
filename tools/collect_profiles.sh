#!/bin/bash
# After `gpurun -- 'bash tools/gpu_profile.sh'`: copy the summaries into profiles/ and add the units of
# work of each capture to the counters file (bench.py scales the counters by them).
set -e
cd "$(dirname "$0")/.."
python - <<'PY'
import json
units = {"pair_kernel_c2": ("pairs", 4999950000), "pair_kernel_c3": ("pairs", 499999500000),
         "search_kriging": ("targets", 262144), "solve_chol_kriging": ("targets", 262144),
         "search_kriging_c5": ("targets", 125000), "solve_ldl_kriging_c5": ("targets", 125000),
         "solve_lu_kriging_c5": ("targets", 125000)}
new = json.load(open("gpurun_out/r02_kernel_counters.json"))
for k, v in new.items():
    v["unit_of_work"], v["units_in_capture"] = units[k]
json.dump(new, open("profiles/r02_kernel_counters.json", "w"), indent=1, sort_keys=True)
print({k: v["build"] for k, v in new.items()})
PY
for n in pair_c3 pair_c2 chol_c4 search_c4 ldl_c5 lu_c5 search_c5; do cp gpurun_out/r02_ncu_${n}_summary.txt profiles/; done
for f in r02_ncu_pair_c3_lines.txt r02_ncu_chol_c4_lines.txt r02_ncu_lu_c5_lines.txt r02_ncu_ldl_c5_lines.txt \
         r02_ncu_search_c4_lines.txt r02_ncu_lu_c5_regions.txt r02_launches_bench.csv; do cp gpurun_out/$f profiles/; done
python tools/launch_summary.py profiles/r02_launches_bench.csv > profiles/r02_launch_summary.txt
python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt
head -c 16 stif_b200/libstif_b200.so.stamp; echo
