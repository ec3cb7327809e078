# round-2 tile-shape A/B (variants from scripts/build_variant.sh)
bash scripts/ab_libs.sh 'kw' topoflow36_b200/libtfb_b200.so topoflow36_b200/_variants/libtfb_none_r22.so topoflow36_b200/_variants/libtfb_none_r24.so topoflow36_b200/_variants/libtfb_none_r26.so topoflow36_b200/_variants/libtfb_none_r30.so
bash scripts/ab_libs.sh 'kw_ga' topoflow36_b200/libtfb_b200.so topoflow36_b200/_variants/libtfb_ga_r20.so topoflow36_b200/_variants/libtfb_ga_r22.so
