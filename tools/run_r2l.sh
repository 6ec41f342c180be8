V=/root/repo/-reinforcement-learning-for-dynamic-uav-energy-efficient-path-planning-in-iot-sensor-networks._b200/variants
python -m pytest tests/test_gpu_parity.py -m gpu -q -k config3_full 2>&1 | tail -12 | cut -c1-2500 > gpurun_out/r2l_tests.log
STEPS=300 BENCH_FLAGS="--sustained-s 0" tools/ab.sh gpurun_out/r2l_ab.txt "X=1" "UAVSIM_LIB=$V/libuavsim_w7b3.so" "UAVSIM_LIB=$V/libuavsim_w21.so" "UAVSIM_LIB=$V/libuavsim_w11b2.so"
cat gpurun_out/r2l_tests.log
