#!/bin/bash
# Regenerates the committed golden fixtures by running the UNMODIFIED reference
# (oracle/_ref/ref_driver, built by `make -C oracle ref` from /root/reference).
# Each .oshd.gz holds one refine_by_size pass: input mesh, derived adjacencies,
# stage intermediates and the output mesh (see oracle/ref_driver.cpp).
set -e
cd "$(dirname "$0")"
R=../../oracle/_ref/ref_driver
T=$(mktemp -d)
#        dim n metric passes prefix        [minq]
$R refine 2  2 0 2 $T/d2n2m0 > /dev/null      # SURVEY.md Appendix D worked example
$R refine 2  6 1 2 $T/d2n6m1 > /dev/null      # 2-D anisotropic (rotated tanh layer)
$R refine 2  6 2 3 $T/d2n6m2 > /dev/null
$R refine 3  2 0 5 $T/d3n2m0 > /dev/null      # 3-D isotropic, all 4 doubling passes + the no-op 5th call
$R refine 3  3 0 2 $T/d3n3m0 > /dev/null      # non-power-of-two grid (inexact coordinates)
$R refine 3  4 2 2 $T/d3n4m2 > /dev/null      # 3-D anisotropic, distinct eigenvalues
$R refine 3  4 1 1 $T/d3n4m1 > /dev/null      # 3-D anisotropic, repeated eigenvalues
$R refine 3  4 3 3 $T/d3n4m3 0.47 > /dev/null # corner_test.cpp metric + min_quality_allowed=0.47
# user fields through TransferOpts::type_map: temperature LINEAR_INTERP, aux_metric METRIC, mat_id INHERIT, rho DENSITY
$R refine 3  3 0 2 $T/d3n3m0x 0 1 1 > /dev/null
$R refine 2  6 1 2 $T/d2n6m1x 0 1 1 > /dev/null
# Mesh::balance's recursive inertial bisection (element -> part) by the reference's inertia::mark_bisection
#      dim nx ny nz nparts
$R rib 3  4  4  4  4 $T/rib_d3n4p4.oshd
$R rib 3  8  4  4  8 $T/rib_d3n8x4x4p8.oshd
$R rib 2 12  8  0  4 $T/rib_d2n12x8p4.oshd
$R rib 3 10  7  5  8 $T/rib_d3n10x7x5p8.oshd
for f in $T/*.oshd; do gzip -9 -n -c $f > $(basename $f).gz; done
rm -rf $T
ls -la *.gz | awk '{s+=$5} END {print NR, "fixtures,", s, "bytes"}'
