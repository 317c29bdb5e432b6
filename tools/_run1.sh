for i in 1 2; do
python tools/time_modes.py --modes fast --weeks 128 2>/dev/null | cut -c1-400
cp xgc_msm_b200/libxgc.so /tmp/new.so; cp xgc_msm_b200/libxgc_alt.so xgc_msm_b200/libxgc.so
python tools/time_modes.py --modes fast --weeks 128 2>/dev/null | head -1 | cut -c1-120
cp /tmp/new.so xgc_msm_b200/libxgc.so
done
