for v in lib_t64; do
  echo "== $v"; MINIRENDER_B200_LIB=$PWD/minirender_b200/lib/$v.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1; for sc in sphere bench cloud; do MINIRENDER_B200_LIB=$PWD/minirender_b200/lib/$v.so python tools/stage_probe.py $sc 0 | grep flags; done
done
