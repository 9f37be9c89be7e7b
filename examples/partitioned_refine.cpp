// A C++ host of the partitioned refine loop, over the C ABI only (no Python, no torch):
//   oshb_build_box -> oshb_dist_distribute (this rank's part + halo) -> loop { oshb_dist_refine_by_size;
//   oshb_dist_reghost when the halo is used up } -> compare with the serial loop.
// It is what an MPI rank of a simulation code does: every rank runs this with its own (rank, size) and a
// communicator whose collectives are MPI_Allreduce / MPI_Allgather / MPI_Alltoallv (oshb_comm_create_callbacks) or
// NCCL (oshb_comm_create_nccl with the id of rank 0 broadcast by the host's own means). Compiled and tested here as a
// single rank -- the product build over NCCL (a world of one), the emulation build over callbacks --; the multi-rank
// runs of the same entry points are tests/test_dist.py and bench.py --gpus N.
//
//   g++ -std=c++17 -O2 examples/partitioned_refine.cpp -Iinclude -Lomega_h_b200/lib -loshb
#include <oshb.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>

static void check(int rc) {
  if (rc != 0) throw std::runtime_error(oshb_last_error());
}

// a world of one rank: every collective is a copy. Buffers are DEVICE pointers, staged through the host here exactly
// where an MPI host without CUDA-aware MPI would call MPI_Allgather / MPI_Alltoallv on the staged bytes
static int staged_copy(void* d_dst, const void* d_src, size_t bytes) {
  if (bytes == 0) return 0;
  std::vector<char> stage(bytes);
  if (oshb_d2h(stage.data(), d_src, bytes) != 0) return 1;
  return oshb_h2d(d_dst, stage.data(), bytes);
}
static int one_allreduce(void*, int32_t*, int) { return 0; }
static int one_allgather(void*, const int64_t* d_send, int n, int64_t* d_recv) {
  return staged_copy(d_recv, d_send, size_t(n) * sizeof(int64_t));
}
static int one_alltoallv(void*, const void* d_send, const int64_t* sc, void* d_recv, const int64_t* rc, int eb) {
  if (sc[0] != rc[0]) return 1;
  return staged_copy(d_recv, d_send, size_t(sc[0]) * size_t(eb));
}

int main(int argc, char** argv) {
  int const n = (argc > 1) ? atoi(argv[1]) : 8;
  int const halo = (argc > 2) ? atoi(argv[2]) : 2;
  try {
    check(oshb_init(0));
    int const rank = 0, size = 1;
    oshb_comm* comm = nullptr;
    if (oshb_is_emulation()) {
      oshb_comm_callbacks cb;
      cb.user = nullptr;
      cb.allreduce_max_i32 = one_allreduce;
      cb.allgather_i64 = one_allgather;
      cb.alltoallv = one_alltoallv;
      check(oshb_comm_create_callbacks(rank, size, &cb, /*sync_first=*/1, &comm));
    } else {
      char id[128];
      check(oshb_comm_nccl_unique_id(id));  // rank 0; an MPI host broadcasts these 128 bytes
      check(oshb_comm_create_nccl(rank, size, id, &comm));
    }
    // every rank builds (or reads) the same mesh and keeps its part
    oshb_mesh* full = nullptr;
    check(oshb_build_box(1.0, 1.0, 1.0, n, n, n, &full));
    int32_t nverts = 0;
    check(oshb_mesh_nents(full, 0, &nverts));
    double const h = 1.0 / n / 2.0;
    std::vector<double> metric(size_t(nverts), 1.0 / (h * h));
    check(oshb_mesh_add_tag(full, 0, "metric", /*F64*/ 3, 1, metric.data(), /*host=*/1, /*internal=*/0));
    oshb_mesh* part = nullptr;
    check(oshb_dist_distribute(full, rank, size, halo, /*parting: ranges*/ 0, &part));
    int64_t nglobal[4] = {0, 0, 0, 0};
    for (int d = 0; d < 4; ++d) {
      int32_t c = 0;
      check(oshb_mesh_nents(full, d, &c));
      nglobal[d] = c;
    }
    oshb_adapt_opts opts;
    check(oshb_adapt_opts_init(3, &opts));
    // the serial loop on the whole mesh, for comparison
    int serial_passes = 0;
    for (int did = 1; did;) {
      check(oshb_refine_by_size(full, &opts, &did));
      serial_passes += did;
    }
    // the partitioned loop
    int passes = 0, total = 0, reghosts = 0, result = 1;
    while (true) {
      check(oshb_dist_refine_by_size(part, comm, &opts, halo, &passes, nglobal, &result, nullptr));
      if (result == 0) break;
      if (result == 2) {
        check(oshb_dist_reghost(part, comm, halo));
        passes = 0;
        ++reghosts;
        continue;
      }
      ++total;
      printf("pass %d: %lld tets\n", total, (long long)nglobal[3]);
    }
    int32_t serial_tets = 0, part_tets = 0;
    check(oshb_mesh_nents(full, 3, &serial_tets));
    check(oshb_mesh_nents(part, 3, &part_tets));
    printf("partitioned: %d passes, %d re-ghostings, %lld tets; serial: %d passes, %d tets\n", total, reghosts,
        (long long)nglobal[3], serial_passes, serial_tets);
    bool const ok = (total == serial_passes) && (nglobal[3] == serial_tets) && (part_tets == serial_tets);
    oshb_mesh_destroy(part);
    oshb_mesh_destroy(full);
    oshb_comm_destroy(comm);
    printf(ok ? "PARTITIONED_OK\n" : "PARTITIONED_MISMATCH\n");
    return ok ? 0 : 2;
  } catch (std::exception const& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}
