#!/bin/sh
# OPTIONAL GPU BASELINE (bench evidence only, never on the product path): the UNMODIFIED reference, built as its
# authors build it for NVIDIA GPUs — vendored Kokkos 5.1.0 with the CUDA backend for Blackwell (generic sm_100, the
# arch Kokkos knows: external/kokkos/cmake/kokkos_arch.cmake:96) + src/main.cpp — following SURVEY.md Appendix A.
# Everything is built under $B (default /tmp/e2d_ref_cuda); only the executable is kept, in baseline/_ref/
# (git-ignored, travels to the GPU box).  The reference's top-level CMakeLists needs Fortran + HWLOC, hence the
# 12-line out-of-tree CMakeLists written below; no reference source is copied or modified.
set -e
REF=${REF:-/root/reference}
B=${B:-/tmp/e2d_ref_cuda}
HERE=$(cd "$(dirname "$0")" && pwd)
[ -f "$REF/src/main.cpp" ] || { echo "reference tree not present: keeping any prebuilt baseline/_ref/"; exit 0; }
export PATH=/usr/local/cuda/bin:$PATH
export NVCC_WRAPPER_DEFAULT_COMPILER=/usr/bin/g++
mkdir -p "$B"
CXX=$REF/external/kokkos/bin/nvcc_wrapper CC=/usr/bin/gcc cmake -S "$REF/external/kokkos" -B "$B/kokkos" \
  -DCMAKE_BUILD_TYPE=Release -DCMAKE_CXX_STANDARD=20 -DKokkos_ENABLE_CUDA=ON -DKokkos_ENABLE_SERIAL=ON \
  -DKokkos_ENABLE_CUDA_CONSTEXPR=ON -DKokkos_ARCH_BLACKWELL100=ON -DKokkos_ENABLE_HWLOC=OFF -DKokkos_ENABLE_TESTS=OFF \
  -DCMAKE_INSTALL_PREFIX="$B/kokkos-install" > "$B/kokkos-configure.log" 2>&1
make -C "$B/kokkos" -j8 install > "$B/kokkos-build.log" 2>&1
mkdir -p "$B/app"
cat > "$B/app/CMakeLists.txt" <<EOF
cmake_minimum_required(VERSION 3.20)
project(euler2d_ref_cuda LANGUAGES CXX)
set(CMAKE_CXX_STANDARD 20)
set(CMAKE_CXX_EXTENSIONS OFF)
find_package(ZLIB REQUIRED)
find_package(Kokkos 5.1.0 CONFIG REQUIRED)
add_executable(euler2d_kokkos_cuda $REF/config/inih/ini.cpp $REF/config/inih/INIReader.cpp $REF/config/ConfigMap.cpp
               $REF/src/HydroParams.cpp $REF/src/SimpleTimer.cpp $REF/src/cnpy/cnpy.cpp $REF/src/main.cpp)
target_compile_definitions(euler2d_kokkos_cuda PRIVATE USE_DOUBLE)
target_include_directories(euler2d_kokkos_cuda PUBLIC $REF $REF/src)
target_link_libraries(euler2d_kokkos_cuda Kokkos::kokkos ZLIB::ZLIB)
EOF
CXX=$REF/external/kokkos/bin/nvcc_wrapper cmake -S "$B/app" -B "$B/app/build" -DCMAKE_BUILD_TYPE=Release \
  -DKokkos_DIR="$B/kokkos-install/lib/cmake/Kokkos" > "$B/app-configure.log" 2>&1
make -C "$B/app/build" -j8 > "$B/app-build.log" 2>&1
mkdir -p "$HERE/_ref"
cp "$B/app/build/euler2d_kokkos_cuda" "$HERE/_ref/"
echo "built $HERE/_ref/euler2d_kokkos_cuda"
