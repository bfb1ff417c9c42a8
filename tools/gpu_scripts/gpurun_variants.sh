for v in $VARIANTS; do
  echo "== $v"; for sc in $SCENES; do MINIRENDER_B200_LIB=$PWD/minirender_b200/lib/$v.so python tools/stage_probe.py $sc 0 | grep flags; done
done
