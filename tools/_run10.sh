python -m pytest tests/test_gpu_parity.py -x -q -k "taps128 or golden or auto or device_pointers or batch_api or scale_images" 2>&1 | tail -2
for j in "3840 2160 1280 720 0 0 1" "3840 2160 1280 720 8 8 1" "3840 2160 1280 720 4 4 0" "3840 2160 800 450 1 5 1" "4000 3000 640 480 0 0 1"; do
  python tools/time_job.py $j 8 --align
done
