#!/bin/bash
# Regenerates the evidence behind the RRT_FLAG_FMAD contract: the SASS of the reference's own kernel, built the way
# oracle/Makefile builds oracle/_ref/libref_cuda.so (nvcc -O3 sm_100a, nvcc's default -fmad=true), SPIN_A = 0.99.
# Needs the reference tree; writes only under /tmp.  Usage: tests/tools/dump_ref_sass.sh [/root/reference]
#
# Where to look in /tmp/rrt_ref_sass/ref.sass (nvcc 12.9):
#   ray setup            the code before the first MUFU.RSQ: uv = x/w, y/h (div), lens: FFMA tu*tu + (tv*tv), FFMA rr*k + 1,
#                        FFMA tu*f + 0.5; direction: FMUL vc*up, FFMA uc*right + ., FADD . + forward; |d|^2 = FFMA chain
#   loop header          FMUL x*x, FMUL z*z, FFMA y*y + x*x, FADD . + z*z   (x*x and z*z are shared with the density code)
#   RK4 stage 1          cross = FFMA a*b - (c*d); L^2 = FFMA chain; FMUL L2*-3; den = (r2*r2)*r; inline div; then
#                        FMUL drag*s and FFMA m*p + (drag*s)      <- first product fused
#   RK4 stages 2-4       ... FMUL p*m and FFMA drag*s + (p*m)      <- second product fused
#   stage updates        FFMA v*hh + p, FFMA hh*k + v;  sums FFMA k3*2 + k4, FFMA k2*2 + ., FADD . + k1;  FFMA sum*h6 + p
#   escape test          FMUL vy*y, FFMA vx*x + ., FFMA vz*z + .
#   noise (called fn)    hash: FMUL (z+K)*y, FFMA (y+K)*x + ., FFMA (x+K)*z + .;  lerp: FADD b-a, FFMA t*(b-a) + a;
#                        fbm: FFMA p*2.05 + 10, FFMA noise*amp + value
#   epilogue             FFMA bg*T + I;  bloom: FMUL g*.7152, FFMA r*.2126 + ., FFMA b*.0722 + .;  vignette: FFMA d*k - 0.8
set -e
REF=${1:-/root/reference}
OUT=/tmp/rrt_ref_sass
mkdir -p $OUT/cfg
sed 's/#define SPIN_A 0.0f /#define SPIN_A 0.99f/' $REF/include/config.h > $OUT/cfg/config.h
grep -q 'SPIN_A 0.99f' $OUT/cfg/config.h
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -I$OUT/cfg -I$REF/include -cubin $REF/src/raymarcher.cu -o $OUT/ref.cubin
cuobjdump -sass $OUT/ref.cubin | grep -E '^\s+/\*[0-9a-f]{4,5}\*/' | sed 's/;.*//' > $OUT/ref.sass
echo "wrote $OUT/ref.sass ($(wc -l < $OUT/ref.sass) instructions)"
grep -c "FFMA" $OUT/ref.sass | sed 's/^/FFMA count: /'
