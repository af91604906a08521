set -u
cp minizero_b200/lib/libmzb200.so /tmp/libmzb200_release.so
MZ_BUILD_EXPERIMENT=1 python -c "import minizero_b200; minizero_b200.build_library(force=True)"
for set in "MZ_CARVEOUT=0" "MZ_CARVEOUT=1" "MZ_CARVEOUT=0" "MZ_CARVEOUT=1"; do
  env KT_CONFIG=2 $set timeout 300 python profiles/kernel_times.py 2>&1 | tail -1
done
env KT_CONFIG=4 MZ_CARVEOUT=1 timeout 300 python profiles/kernel_times.py 2>&1 | tail -1
env KT_CONFIG=4 MZ_CARVEOUT=0 timeout 300 python profiles/kernel_times.py 2>&1 | tail -1
cp /tmp/libmzb200_release.so minizero_b200/lib/libmzb200.so
