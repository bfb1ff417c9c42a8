# stage times of library variants: VARIANTS="libA libB" SCENES="sphere bench"
for v in $VARIANTS; do
  echo "== $v"; for sc in ${SCENES:-sphere}; do MINIRENDER_B200_LIB=$PWD/minirender_b200/lib/$v.so timeout 60 python tools/stage_probe.py $sc 0 2>&1 | grep -E "flags|rror"; done
done
