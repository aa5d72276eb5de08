#!/bin/bash
# Regenerates the round's measurement artefacts on a GPU box (run under gpurun from the repo root):
# bench lines for the five BASELINE configurations (device-resident figure) and the full default line,
# the reference arm, the ncu launch list of the default bench command, a graph-level ncu capture of one
# cfg 2 step (16 launches as ONE workload: total DRAM traffic with the launches overlapping), one
# `ncu --set full` capture per configuration's kernel with its per-instruction execution counts, the
# type-pair matrices, the cliff table and the bandwidth probes.  Everything lands in gpurun_out/;
# copy what should be judged to profiles/.
R=${1:-r02}
O=gpurun_out
mkdir -p $O
python bench.py 2>/dev/null | tail -1 > $O/${R}_bench_cfg2.json
for c in cfg1 cfg3 cfg4 cfg5; do
  python bench.py --config $c --no-multi-gpu 2>/dev/null | tail -1 > $O/${R}_bench_$c.json
done
python bench.py --config cfg4 --batched --quick 2>/dev/null | tail -1 > $O/${R}_bench_cfg4_batched.json
python bench.py --impl reference 2>/dev/null | tail -1 > $O/${R}_bench_reference_cfg2.json
./tools/hbm_probe 2048 10 > $O/${R}_hbm_probe.json 2>/dev/null
./tools/pcie_probe 256 8 > $O/${R}_pcie_probe_n1.json 2>/dev/null
./tools/call_overhead smolscale_b200/libsmolscale_cuda.so > $O/${R}_call_overhead.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 1 --passes 2 --quick > /dev/null 2>&1
ncu --graph-profiling graph --profile-from-start off --clock-control none \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum \
    -c 3 --csv --log-file $O/${R}_ncu_graph_cfg2.csv python tools/graph_step.py cfg2 4 > /dev/null 2>&1
prof () { # name, kernel regex, command: summary + per-instruction counts; the report itself is dropped (gpurun_out is capped at 64 MiB)
  ncu --set full --clock-control none --import-source on -k regex:$2 -s ${4:-2} -c 1 -f -o $O/${R}_full_$1 $3 > /dev/null 2>&1
  python tools/ncu_summary.py $O/${R}_full_$1.ncu-rep > $O/${R}_ncu_full_$1.txt 2>&1
  python tools/ncu_sass_hot.py $O/${R}_full_$1.ncu-rep 1 > $O/${R}_ncu_sass_hot_$1.txt 2>&1
  python tools/ncu_sass_hot.py $O/${R}_full_$1.ncu-rep 0 > $O/${R}_ncu_sass_all_$1.txt 2>&1
  rm -f $O/${R}_full_$1.ncu-rep
}
prof cfg1 smol_half "python tools/one_conv.py 1920 1080 960 540 0 0 0"
prof cfg2 smol_half "python tools/one_conv.py 3840 2160 1920 1080 1 5 0"
prof cfg3 smol_box "python tools/one_conv.py 7680 4320 800 450 0 0 1"
prof cfg4 smol_magb "python tools/one_conv.py 1024 768 4096 3072 8 8 0"
prof cfg5 smol_half "python tools/graph_step.py cfg5 2" 1
prof conv_uu smol_taps0w "python tools/one_conv.py 3840 2160 3839 2159 4 4 0"
python tools/ncu_sections.py $O/${R}_ncu_sass_all_cfg3.txt > $O/${R}_ncu_sections_cfg3.txt 2>&1
for g in "3840x2160:3839x2159 0 4k_1to1" "3840x2160:3839x2159 1 4k_1to1_srgb" "3840x2160:1280x720 0 4k_3x" "3840x2160:1280x720 1 4k_3x_srgb" "7680x4320:800x450 0 8k_box" "7680x4320:800x450 1 8k_box_srgb" "1920x1080:3840x2160 0 up2x"; do
  set -- $g
  python tools/bench_conv.py --geom $1 --srgb $2 --json $O/${R}_conv_$3.json > /dev/null 2>&1
done
python tools/cliff_table.py > $O/${R}_cliff_table.json 2>/dev/null
ls -la $O | tail -40
