timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "one_halving or random_matrix or magb" 2>&1 | tail -3
python tools/bench_conv.py --geom 3840x2160:1280x720 --srgb 0 --json gpurun_out/r02_conv_4k_3x.json > /dev/null 2>&1
python - <<'P'
import json,statistics
d=json.load(open('gpurun_out/r02_conv_4k_3x.json'))
us=[r['us'] for r in d['rows']]
print('conv 4k_3x', round(min(us),1), round(statistics.median(us),1), round(max(us),1), all(r['ok'] for r in d['rows']))
P
python tools/time_job.py 3840 2160 1280 720 0 0 0 16
python tools/time_job.py 3840 2160 1600 900 0 0 0 16
python tools/time_job.py 1920 1080 854 480 0 0 0 16
python tools/time_job.py 3840 2160 1280 720 8 8 0 16
timeout 250 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "one_halving or magb_word" 2>&1 | tail -4
