for v in libminirender_b200 lib_s5 lib_s6 lib_r5 lib_r6 lib_r3; do
  echo "== $v"; MINIRENDER_B200_LIB=$PWD/minirender_b200/lib/$v.so python tools/stage_probe.py sphere 0 | grep flags
  MINIRENDER_B200_LIB=$PWD/minirender_b200/lib/$v.so python tools/stage_probe.py bench 0 | grep flags
done
