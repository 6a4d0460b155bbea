cd $GRAFT_REPO_ROOT
mkdir -p /tmp/objs; for f in tools/_var/objs/*.obj; do cp $f /tmp/objs/$(basename $f .obj).o; done
for f in tools/_var/scan_*.obj; do
echo "== $f"
cp $f /tmp/scan_var.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o pdl_b200/lib/libpdlb200.so /tmp/objs/*.o /tmp/scan_var.o -Xlinker --exclude-libs,ALL || continue
timeout 300 python tools/scan_bench.py 2>&1 | grep "float\[268435456\]\|double\[134217728\]" | cut -c1-200
done
