#!/bin/bash
# Regenerates the round's measurement artefacts on a GPU box (run under gpurun from the repo root):
# bench lines for the five BASELINE configurations, the reference arm, the ncu launch list of the
# default bench command, one `ncu --set full` capture per configuration's kernel, the type-pair
# matrices and the bandwidth probe.  Everything lands in gpurun_out/; summaries are made from the
# reports on the box with tools/ncu_summary.py / ncu_sass_hot.py; copy what should be judged to profiles/.
R=${1:-r01}
O=gpurun_out
mkdir -p $O
for c in cfg1 cfg2 cfg3 cfg4 cfg5; do
  python bench.py --config $c 2>/dev/null | tail -1 > $O/${R}_bench_$c.json
done
python bench.py --config cfg4 --batched --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > $O/${R}_bench_cfg4_batched.json
python bench.py --impl reference 2>/dev/null | tail -1 > $O/${R}_bench_reference_cfg2.json
python tools/bw_probe.py > $O/${R}_bw_probe.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
prof () { # name, kernel regex, one_conv args: summary + hottest SASS lines; the report itself is dropped (gpurun_out is capped at 64 MiB)
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o $O/${R}_full_$1 python tools/one_conv.py $3 > /dev/null 2>&1
  python tools/ncu_summary.py $O/${R}_full_$1.ncu-rep > $O/${R}_ncu_full_$1.txt 2>&1
  python tools/ncu_sass_hot.py $O/${R}_full_$1.ncu-rep 1 > $O/${R}_ncu_sass_hot_$1.txt 2>&1
  rm -f $O/${R}_full_$1.ncu-rep
}
prof cfg1 smol_half "1920 1080 960 540 0 0 0"
prof cfg2 smol_half "3840 2160 1920 1080 1 5 0"
prof cfg3 smol_box "7680 4320 800 450 0 0 1"
prof cfg4 smol_magb "1024 768 4096 3072 8 8 0"
prof cfg5 smol_half "2048 2048 256 256 2 2 0"
for g in "3840x2160:3839x2159 0 4k_1to1" "3840x2160:1280x720 0 4k_3x" "3840x2160:1280x720 1 4k_3x_srgb" "7680x4320:800x450 0 8k_box" "7680x4320:800x450 1 8k_box_srgb" "1920x1080:3840x2160 0 up2x"; do
  set -- $g
  python tools/bench_conv.py --geom $1 --srgb $2 --json $O/${R}_conv_$3.json > /dev/null 2>&1
done
ls -la $O | tail -30
