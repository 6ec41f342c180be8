V=/root/repo/-reinforcement-learning-for-dynamic-uav-energy-efficient-path-planning-in-iot-sensor-networks._b200/variants
STEPS=300 BENCH_FLAGS="--sustained-s 0" tools/ab.sh gpurun_out/r2q_ab.txt "X=1" "UAVSIM_LIB=$V/libuavsim_r1all.so" "UAVSIM_LIB=$V/libuavsim_f32rew.so" "UAVSIM_LIB=$V/libuavsim_onepass.so" "UAVSIM_LIB=$V/libuavsim_adrimm.so" "X=2"
