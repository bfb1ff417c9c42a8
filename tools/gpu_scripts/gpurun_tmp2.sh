python tools/stage_probe.py cloud 0 | grep flags
for n in 4 8 16 32 64; do echo "setup ctas/sm $n"; MR_SETUP_CTAS_PER_SM=$n python tools/stage_probe.py cloud 0 | grep flags; done
MR_NO_CLUSTER_CULL=1 python tools/stage_probe.py cloud 0 | grep flags
