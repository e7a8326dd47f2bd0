echo default; python tools/gpu_time.py p3_t09.cli 3840 2160 16 3 1 | tail -1
for f in build/libdrt_*.so; do echo $f; DRT_LIB=$f python tools/gpu_time.py p3_t09.cli 3840 2160 16 3 1 | tail -1; done
