#!/bin/sh
# OPTIONAL CPU BASELINE (bench evidence only, never on the product path): the UNMODIFIED reference, built as its
# authors build it for CPUs — vendored Kokkos 5.1.0 with the OpenMP backend + src/main.cpp — following SURVEY.md
# Appendix A.  bench.py reports it beside the oracle/_ref number (the same sources on the test shim's loop runner).
# Everything is built under $B (default /tmp/e2d_ref_omp); only the executable is kept, in baseline/_ref/
# (git-ignored, travels to the GPU box).  The reference's top-level CMakeLists needs Fortran + HWLOC, hence the
# 12-line out-of-tree CMakeLists written below; no reference source is copied or modified.
set -e
REF=${REF:-/root/reference}
B=${B:-/tmp/e2d_ref_omp}
HERE=$(cd "$(dirname "$0")" && pwd)
[ -f "$REF/src/main.cpp" ] || { echo "reference tree not present: keeping any prebuilt baseline/_ref/"; exit 0; }
mkdir -p "$B"
CXX=/usr/bin/g++ CC=/usr/bin/gcc cmake -S "$REF/external/kokkos" -B "$B/kokkos" \
  -DCMAKE_BUILD_TYPE=Release -DCMAKE_CXX_STANDARD=20 -DKokkos_ENABLE_OPENMP=ON -DKokkos_ENABLE_SERIAL=ON \
  -DKokkos_ENABLE_HWLOC=OFF -DKokkos_ENABLE_TESTS=OFF \
  -DCMAKE_INSTALL_PREFIX="$B/kokkos-install" > "$B/kokkos-configure.log" 2>&1
make -C "$B/kokkos" -j8 install > "$B/kokkos-build.log" 2>&1
mkdir -p "$B/app"
cat > "$B/app/CMakeLists.txt" <<EOF
cmake_minimum_required(VERSION 3.20)
project(euler2d_ref_omp LANGUAGES CXX)
set(CMAKE_CXX_STANDARD 20)
set(CMAKE_CXX_EXTENSIONS OFF)
find_package(ZLIB REQUIRED)
find_package(Kokkos 5.1.0 CONFIG REQUIRED)
add_executable(euler2d_kokkos_omp $REF/config/inih/ini.cpp $REF/config/inih/INIReader.cpp $REF/config/ConfigMap.cpp
               $REF/src/HydroParams.cpp $REF/src/SimpleTimer.cpp $REF/src/cnpy/cnpy.cpp $REF/src/main.cpp)
target_compile_definitions(euler2d_kokkos_omp PRIVATE USE_DOUBLE)
target_include_directories(euler2d_kokkos_omp PUBLIC $REF $REF/src)
target_link_libraries(euler2d_kokkos_omp Kokkos::kokkos ZLIB::ZLIB)
# the oracle's driver of the reference (oracle/ref/ref_dump.cpp: raw state dumps, loop timer) on the REAL Kokkos/OpenMP
# runtime instead of oracle/kokkos_shim: installed as oracle/_ref/ref_dump_kokkos, which oracle.ref_binary() prefers
add_executable(ref_dump_kokkos $REF/config/inih/ini.cpp $REF/config/inih/INIReader.cpp $REF/config/ConfigMap.cpp
               $REF/src/HydroParams.cpp $REF/src/SimpleTimer.cpp $REF/src/cnpy/cnpy.cpp $HERE/../oracle/ref/ref_dump.cpp)
target_compile_definitions(ref_dump_kokkos PRIVATE USE_DOUBLE)
target_include_directories(ref_dump_kokkos PUBLIC $REF $REF/src)
target_link_libraries(ref_dump_kokkos Kokkos::kokkos ZLIB::ZLIB)
EOF
CXX=/usr/bin/g++ cmake -S "$B/app" -B "$B/app/build" -DCMAKE_BUILD_TYPE=Release \
  -DKokkos_DIR="$B/kokkos-install/lib/cmake/Kokkos" > "$B/app-configure.log" 2>&1
make -C "$B/app/build" -j8 > "$B/app-build.log" 2>&1
mkdir -p "$HERE/_ref"
cp "$B/app/build/euler2d_kokkos_omp" "$HERE/_ref/"
mkdir -p "$HERE/../oracle/_ref"
cp "$B/app/build/ref_dump_kokkos" "$HERE/../oracle/_ref/"
echo "built $HERE/_ref/euler2d_kokkos_omp and oracle/_ref/ref_dump_kokkos"
