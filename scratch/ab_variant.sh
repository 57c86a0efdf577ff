#!/bin/bash
# A/B an experimental build of the CUDA library against the product library on one B200.
#   here (CPU):   python -c "from sea_ice_drift_b200 import _build; _build.build(force=True, out='scratch/ab/libsid_variant.so', extra=['-DSID_IMMA_KSTD'])"
#                 cp sea_ice_drift_b200/libsid_b200.so scratch/ab/libsid_base.so
#   then:         gpurun --timeout 600 -- 'bash scratch/ab_variant.sh'
# Times both on cfg2 (scratch/time_variants.py also compares each library's IMMA result with its own dp4a path on all
# points), runs the FULL -m gpu suite on the variant, and puts the product library back.  *.so files are git-ignored but
# travel with the gpurun snapshot.
set -u
[ -f scratch/ab/libsid_variant.so ] && [ -f scratch/ab/libsid_base.so ] || { echo "build scratch/ab/libsid_{base,variant}.so first"; exit 2; }
for v in base variant base variant; do
  cp scratch/ab/libsid_$v.so sea_ice_drift_b200/libsid_b200.so
  echo "== $v"; timeout 120 python scratch/time_variants.py cfg2 2>&1 | sed -n 2,4p
done
cp scratch/ab/libsid_variant.so sea_ice_drift_b200/libsid_b200.so
echo "== full GPU suite on the variant"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
cp scratch/ab/libsid_base.so sea_ice_drift_b200/libsid_b200.so
