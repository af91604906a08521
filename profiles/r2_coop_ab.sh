set -u
KT_CONFIG=2 timeout 300 python profiles/kernel_times.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_think.py tests/test_gpu_atari.py -x -q 2>&1 | tail -3
cp minizero_b200/lib/libmzb200.so /tmp/libmzb200_release.so
MZ_BUILD_EXPERIMENT=1 python -c "import minizero_b200; minizero_b200.build_library(force=True)"
for set in "MZ_TOWER_COOP=1" "MZ_TOWER_COOP=0"; do
  env KT_CONFIG=2 $set timeout 300 python profiles/kernel_times.py 2>&1 | tail -1
  env KT_CONFIG=4 $set timeout 300 python profiles/kernel_times.py 2>&1 | tail -1
done
cp /tmp/libmzb200_release.so minizero_b200/lib/libmzb200.so
