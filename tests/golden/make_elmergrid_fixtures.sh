#!/bin/bash
# Regenerates the ElmerGrid fixtures under tests/golden/elmergrid/ with the REFERENCE's own ElmerGrid
# (built by `make -C oracle ref` from /root/reference/elmergrid/src into oracle/_ref/ElmerGrid).
# cube5.grd is fem/tests/PoissonThreaded/cube.grd with `Reference Density = 0.025` replaced by
# `Element Divisions 1/2/3`; the outputs are the reference's mesh.* and partitioning.N files.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
EG=$HERE/../../oracle/_ref/ElmerGrid
W=$(mktemp -d)
cd $W
sed -e 's/^Reference Density = 0.025/Element Divisions 1 = 5\nElement Divisions 2 = 4\nElement Divisions 3 = 3/' \
    /root/reference/fem/tests/PoissonThreaded/cube.grd > cube5.grd
$EG 1 2 cube5 > /dev/null
$EG 1 2 cube5 -partdual -metiskway 3 > /dev/null
mkdir -p $HERE/elmergrid
cp cube5.grd $HERE/elmergrid/
cp cube5/mesh.header cube5/mesh.nodes cube5/mesh.elements cube5/mesh.boundary $HERE/elmergrid/
mkdir -p $HERE/elmergrid/partitioning.3
cp cube5/partitioning.3/part.* $HERE/elmergrid/partitioning.3/
rm -rf $W

# fem/tests/CoordinateScaling: the 20 x 20 quad mesh of square.grd, as that test's runtest.cmake makes it (`ElmerGrid 1 2 square`)
W=$(mktemp -d)
cd $W
cp /root/reference/fem/tests/CoordinateScaling/square.grd .
$EG 1 2 square > /dev/null
mkdir -p $HERE/coordinatescaling
cp square.grd square/mesh.header square/mesh.nodes square/mesh.elements square/mesh.boundary $HERE/coordinatescaling/
rm -rf $W
# fem/tests/ElmerGridExtrudeMaterial: only the .grd is kept; tests/extrudematerial_case.py runs ElmerGrid on it at test time
mkdir -p $HERE/extrudematerial
cp /root/reference/fem/tests/ElmerGridExtrudeMaterial/cubes.grd $HERE/extrudematerial/
