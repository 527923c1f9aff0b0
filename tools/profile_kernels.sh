#!/bin/bash
# ncu --set full captures (one launch each) of the kernels that profiles/ did not cover in round 1; run under gpurun:
#   gpurun -- 'bash tools/profile_kernels.sh'    -> gpurun_out/r02_*.ncu-rep
set -u
mkdir -p gpurun_out
cap() {  # name, kernel regex, run_op args...
  local name=$1 rx=$2; shift 2
  python tools/run_op.py "$@" > /dev/null 2>&1 || { echo "run_op $* failed"; return; }
  ncu --set full --clock-control none --import-source on -k "regex:$rx" -c 1 -f -o gpurun_out/r02_$name python tools/run_op.py "$@" > gpurun_out/ncu_$name.log 2>&1
  tail -n 1 gpurun_out/ncu_$name.log
}
cap bwrf32f_r5_float bwrf32f_tiled bwrf32f_r5 kinect f32
cap bwrf16u_r5_int bwrf32f_tiled bwrf16u_r5 kinect u16
cap bwrf16u_r9_int bwrf32f_tiled bwrf16u_r9 kinect u16
cap bwrf8uc3_r5 bwrf8u_c3 bwrf8uc3_r5 kinect u8
cap bwrf8u_r5 bwrf8u_h2 bwrf8u_r5 kinect u8
cap median_k11 bisect median_k11 kinect u8
cap jpeg_frame jpeg_frame jpeg kinect u8
cap depth32f_range bwrf32f_tiled depth32f kinect u8
cap reproject reproject reproject kinect u8
cap fused13 brf_rank fused13 kinect u8
cap brf13_u8 brf_rank brf13 kinect u8
cap brf13_s16 brf_rank brf13 kinect s16
cap bwrf32fc3_r5 bwrf32f_c3 bwrf32fc3_r5 kinect f32
