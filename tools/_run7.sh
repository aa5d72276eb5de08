python -m pytest tests/test_gpu_parity.py -x -q -k "random_matrix or golden or device_pointers or batch_api" 2>&1 | tail -2
for j in "3840 2160 1280 720 0 0 0" "3840 2160 1280 720 8 8 0" "3840 2160 800 450 1 5 0" "3840 2160 1600 900 0 0 0" "3840 2160 1500 1500 0 0 0"; do
  python tools/time_job.py $j 8 --align
done
