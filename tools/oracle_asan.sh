#!/bin/bash
# The CPU oracle (oracle/mas_oracle.c, test infrastructure) under AddressSanitizer + UBSan over the edge-case meshes.
# It is the yardstick of every GPU parity test, so it has to be clean itself.   bash tools/oracle_asan.sh
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${TMPDIR:-/tmp}/mas_oracle_asan
mkdir -p "$OUT"
FLAGS="-O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off -std=gnu11 -fPIC -fopenmp -mavx2 -mfma"
gcc $FLAGS -DREAL=float -DPFX=maso_f_ -c "$ROOT/oracle/mas_oracle.c" -o "$OUT/f.o"
gcc $FLAGS -DREAL=double -DPFX=maso_d_ -DORACLE_DOUBLE -c "$ROOT/oracle/mas_oracle.c" -o "$OUT/d.o"
gcc -shared -fsanitize=address,undefined -fopenmp -o "$OUT/libmas_oracle.so" "$OUT/f.o" "$OUT/d.o" -lm
cat > "$OUT/run.py" <<PY
import importlib, sys
sys.path.insert(0, "$ROOT"); sys.path.insert(0, "$ROOT/tests")
from oracle import oracle_binding as ob
ob.LIB_PATH = "$OUT/libmas_oracle.so"
S = importlib.import_module("preconditioner-for-cloth-and-deformable-body-simulation_b200").synth
from helpers import make_oracle
def coll(n, seed=7):
    m = S.cloth(n, with_topology=True)
    return S.add_collisions(m, m.nv // 16, m.nv // 16, m.nv // 8, seed=seed)
cases = [lambda: S.cloth(64), lambda: S.cloth(50), lambda: S.cloth(7), lambda: S.cloth(5), lambda: coll(96), lambda: coll(40, 11),
         lambda: S.tet_cube(16, 16, 8), lambda: S.chain(1), lambda: S.chain(32), lambda: S.chain(33), lambda: S.chain(100),
         lambda: S.cloth_with_isolated_vertices(20, 7), lambda: S.dust(6000), lambda: S.dust(31), lambda: S.dust(1025),
         lambda: S.rippled_cloth(64), lambda: S.random_cloud(1500, 5, 3), lambda: S.random_cloud(900, 3, 5),
         lambda: S.stacked_cloth(24, 2), lambda: S.cloth_with_duplicate_edges(24), lambda: S.cloth_rect(96, 40)]
for mk in cases:
    mesh = mk()
    for prec in ("f", "d"):
        o = make_oracle(ob, mesh, prec)
        o.apply(S.residual(mesh.nv))
        o.going_next(); o.coarse_tables(); o.dense_inverse(0); o.mapped_r(); o.mapped_z(); o.fine_connect_mask(); o.sorted_adjacency()
        if o.stencil_num:
            o.stencils()
        o.close()
    print(mesh.name, "clean", flush=True)
PY
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python "$OUT/run.py"
