/* Hand-written stand-in for the header the reference's CMake would generate
 * (cmake/bob.cmake:435 bob_config_header). Serial or OpenMP CPU build, no MPI,
 * no zlib, no Kokkos. TEST INFRASTRUCTURE ONLY (oracle/_ref build). */
#ifndef OMEGA_H_CONFIG_H
#define OMEGA_H_CONFIG_H
#define OMEGA_H_IS_SHARED
#ifdef OSHB_REF_OPENMP
#define OMEGA_H_USE_OPENMP
#endif
#ifdef OSHB_REF_CUDA
#define OMEGA_H_USE_CUDA /* the reference's own (non-Kokkos) CUDA backend: src/Omega_h_for.hpp:16-58 */
#endif
#define OMEGA_H_VERSION_MAJOR 9
#define OMEGA_H_VERSION_MINOR 34
#define OMEGA_H_VERSION_PATCH 13
#define OMEGA_H_SEMVER "9.34.13"
#define OMEGA_H_COMMIT "oracle-ref"
#define OMEGA_H_CXX_FLAGS "-O3 -march=native -ffp-contract=off"
#define OMEGA_H_CMAKE_ARGS ""
#endif
