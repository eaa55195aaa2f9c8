#!/bin/bash
# Turn the outputs of tools/gpu_r2_final.sh (gpurun_out/f_*) into the tracked evidence under profiles/ (round 2).
set -e
cd "$(dirname "$(readlink -f "$0")")/.."
CMD="ncu --set full --import-source on --clock-control none -k regex:detect_kernel -s 3 -c 1 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 --sustained-seconds 0 --no-cli"
python tools/ncu_summary.py gpurun_out/f_full.ncu-rep profiles/r02_final_ncu_summary.json "$CMD" 4096
ncu -i gpurun_out/f_full.ncu-rep --page source --csv > /tmp/r02_src.csv 2>/dev/null
python tools/ncu_opcodes.py /tmp/r02_src.csv 4096 profiles/r02_opcode_mix.json > profiles/r02_final_opcode_mix.txt
python tools/ncu_segments.py /tmp/r02_src.csv 0.4 > profiles/r02_final_segments.txt
cp gpurun_out/f_launches.csv profiles/r02_final_launches.csv
cp gpurun_out/f_sweep.jsonl profiles/r02_sweep.jsonl
cp gpurun_out/f_bench.json profiles/r02_bench.json
cp gpurun_out/f_bench_ref.json profiles/r02_bench_reference_arm.json
# dram traffic per launch of the headline kernel, for bench.py's roofline.traffic
python - <<'PY'
import json
m = json.load(open("profiles/r02_final_ncu_summary.json"))["metrics"]
rd, wr = m["dram__bytes_read.sum"], m["dram__bytes_write.sum"]
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
total = rd["value"] * scale[rd["unit"]] + wr["value"] * scale[wr["unit"]]
json.dump({"dram_bytes_per_launch": total, "source": "profiles/r02_final_ncu_summary.json (dram__bytes_read.sum + dram__bytes_write.sum, one 4096-block launch)"},
          open("profiles/ncu_traffic.json", "w"), indent=1)
print("dram bytes per launch", total)
PY
# SASS excerpt of the headline kernel: the instructions that prove the design (TMA bulk copy, packed FP32, register re-split)
LIB=thrifty_b200/_lib/libthrifty_b200.so
FN=$(cuobjdump --dump-resource-usage $LIB | grep -o "_ZN3thr13detect_kernelILi14ELi512ELb0ELb0ELb0ELi0EEEvNS_12DetectParamsE" | head -1)
cuobjdump -sass -fun $FN $LIB > /tmp/r02_headline.sass
{
  echo "SASS of the headline kernel ($FN) in $LIB, built by thrifty_b200/csrc/Makefile (sm_100a)."
  echo "instruction counts (static):"
  for op in UBLKCP SYNCS FFMA2 FADD2 FMUL2 USETMAXREG LDL STL DFMA LDS STS "BAR.SYNC" "BAR.ARV" UTMALDG UTCHMMA LDTM; do
    printf "  %-10s %s\n" "$op" "$(grep -c "[^A-Z]$op" /tmp/r02_headline.sass)"
  done
  echo "worker part only (after USETMAXREG.TRY_ALLOC):"
  awk '/USETMAXREG.TRY_ALLOC/{f=1} f' /tmp/r02_headline.sass > /tmp/r02_worker.sass
  for op in LDL STL DFMA FFMA2; do printf "  %-10s %s\n" "$op" "$(grep -c "[^A-Z]$op" /tmp/r02_worker.sass)"; done
  echo; echo "register re-split and TMA bulk copies:"
  grep -n "USETMAXREG\|UBLKCP" /tmp/r02_headline.sass | cut -c1-150
  echo; echo "first packed-FP32 butterfly instructions of pass 1:"
  grep -n "FFMA2\|FADD2" /tmp/r02_worker.sass | head -12 | cut -c1-150
} > profiles/r02_headline_sass_excerpt.txt
head -30 profiles/r02_headline_sass_excerpt.txt
