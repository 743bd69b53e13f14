for lib in libdiygym_b200.so libdg_a.so libdg_b.so libdg_c.so; do
echo "== $lib"
DG_LIB=$PWD/diy_gym_b200/$lib timeout 300 python tools/gpu_probe3.py ur_high_5:4:32:0 ur_high_5:4:64:0 ur_high_5:8:32:0 r2d2_maze:8:0:0 r2d2_maze:8:32:0 from_the_readme:8:0:0 ur_high_5:4:64:0:65536 ur_high_5:8:32:0:65536
done
