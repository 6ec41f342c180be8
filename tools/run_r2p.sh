V=/root/repo/-reinforcement-learning-for-dynamic-uav-energy-efficient-path-planning-in-iot-sensor-networks._b200/variants
STEPS=300 BENCH_FLAGS="--sustained-s 0" tools/ab.sh gpurun_out/r2p_ab.txt "X=1" "UAVSIM_LIB=$V/libuavsim_unfold.so" "UAVSIM_LIB=$V/libuavsim_prev.so" "X=2"
