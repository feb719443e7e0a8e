#!/bin/bash
# compute-sanitizer pass over a small invocation of every processor kind (run under gpurun)
cat > /tmp/san.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import mcarray_b200 as mb
from mcarray_b200 import scenes
fs = 16000
xyz4 = scenes.linear_array([0, 0.07, 0.175, 0.21])
xyz16 = scenes.linear_array((np.arange(16) - 7.5) * 0.035)
x4 = scenes.far_field_scene(xyz4, fs, 6000, scenes.azimuth_dirs([0.5]), seed=1).astype(np.float32)
x16 = scenes.far_field_scene(xyz16, fs, 6000, scenes.azimuth_dirs([0.5]), seed=2).astype(np.float32)
x2 = scenes.far_field_scene(scenes.linear_array([0, 0.086]), fs, 6000, scenes.azimuth_dirs([0.5]), seed=3).astype(np.float32)
xyz64 = scenes.planar_array(8, 8, 0.04)
x64 = scenes.far_field_scene(xyz64, fs, 6000, scenes.az_el_dirs(np.array([[0.4]]), np.array([[0.6]])), seed=4).astype(np.float32)
rng = np.random.default_rng(5)
w_fs = rng.standard_normal((7, 16, 257)) + 1j * rng.standard_normal((7, 16, 257))
for p, x in ((mb.SourceSeparationAndLocalisation(fs, xyz4, 1, usePowerFloor=False, max_frames_per_call=32), x4),
             (mb.SourceSeparationAndLocalisation(fs, xyz16, 2, usePowerFloor=True, max_frames_per_call=32), x16),
             (mb.SourceLocalisation(fs, xyz16, 1, usePowerFloor=False, max_frames_per_call=32), x16),
             (mb.FreqGCCBinauralLocalisation(fs, 0.086, usePowerFloor=False, max_frames_per_call=32, frame_size=512), x2),
             (mb.MultibandBinarualLocalisation(fs, 0.086, max_frames_per_call=32), x2),
             (mb.FastBinauralMasking(fs, 0.086, 500, 5000, "RELATIVE", "BOTH", max_frames_per_call=32), x2),
             (mb.TdoaEstimator(fs, 4, 1024, 20, max_frames_per_call=32, emit_curves=True), x4),
             (mb.TdoaEstimator(fs, 4, 256, 9, max_frames_per_call=64), x4),
             (mb.DelayAndSumFan(fs, xyz4, 2048, np.deg2rad(np.arange(-90, 91, 10.0)), max_frames_per_call=8), x4),
             (mb.SrpPhat(fs, xyz16, 512, scenes.az_el_dirs(np.linspace(-3, 3, 50)[:, None], np.array([0.2, 0.9])[None, :]), max_frames_per_call=32), x16),
             (mb.SrpPhat(fs, xyz64, 1024, scenes.az_el_dirs(np.linspace(-3, 3, 40)[:, None], np.array([0.2, 0.9])[None, :]), max_frames_per_call=32), x64),   # srp_tc_kernel, stft_hw<1024>
             (mb.DelayAndSumFan(fs, xyz16, 1024, np.deg2rad(np.arange(-90, 91, 5.0)), max_frames_per_call=32), x16),                                         # ds_fan_tc_kernel
             (mb.FilterAndSumFan(fs, 16, 512, w_fs, max_frames_per_call=32), x16)):
    for pos in range(0, x.shape[1], 2500):
        p.process(x[:, pos:pos + 2500])
    p.synchronize(); p.close()
print("sanitizer workload done")
PY
for tool in memcheck racecheck synccheck; do
  echo "== $tool"; timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py 2>&1 | tail -8
done
