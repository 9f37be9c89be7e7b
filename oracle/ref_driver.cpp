// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
//
// Driver over the UNMODIFIED reference library (oracle/_ref/libomega_h_ref*.so, built by
// oracle/Makefile from the sources where they lie under /root/reference/src). It drives the
// reference's own public API -- build_box (src/Omega_h_build.cpp:136), Mesh::add_tag,
// AdaptOpts (src/Omega_h_adapt.cpp:52-85), refine_by_size (src/Omega_h_refine.cpp:92) -- and
// dumps, per refine pass, the input mesh, the intermediate arrays of the hot path and the
// output mesh into a flat record file ("OSHD1") that tests/ read with numpy.
//
// Modes:
//   refine <dim> <n> <metric> <npasses|-1> <out_prefix> [minq] [askq] [userfields]   per-pass dumps; userfields=1 adds
//          four user tags carried by TransferOpts::type_map: "temperature" (LINEAR_INTERP), "aux_metric" (METRIC),
//          "mat_id" on every dimension (INHERIT), "rho" on elements (DENSITY)
//   time   <dim> <n> <metric> [maxpasses]                              JSON timing line (CPU baseline)
//   timeloops <dim> <n> <metric> <reps> <budget_s>                     the WHOLE loop, repeated on copies of one input
//                                                                      mesh; one JSON line per repetition (bench.py --impl reference)
//   box    <dim> <n> <out>                                             build_box dump only
//   rib    <dim> <nx> <ny> <nz> <nparts> <out>                         element -> part by the reference's inertia::mark_bisection,
//                                                                      applied recursively as recursively_bisect does
//   adjtime <n>                                                        invert_adj / reflect_down timing
//   writeosh <dim> <n> <metric> <npasses> <path.osh> <dump>            binary::write + dump of the same mesh
//   readosh <path.osh> <dump>                                          binary::read + dump
//   diff <a.osh> <b.osh> <tolerance> <floor>                           compare_meshes (what oshdiff runs); prints the result code
// metric: 0 iso h=1/(2n) | 1 tanh layer hx=hy | 2 tanh layer hy=0.7hx | 3 corner_test graded iso
#include <Omega_h_adapt.hpp>
#include <Omega_h_adj.hpp>
#include <Omega_h_array_ops.hpp>
#include <Omega_h_build.hpp>
#include <Omega_h_compare.hpp>
#include <Omega_h_file.hpp>
#include <Omega_h_for.hpp>
#include <Omega_h_indset.hpp>
#include <Omega_h_inertia.hpp>
#include <Omega_h_map.hpp>
#include <Omega_h_mesh.hpp>
#include <Omega_h_metric.hpp>
#include <Omega_h_modify.hpp>
#include <Omega_h_refine.hpp>
#include <Omega_h_refine_qualities.hpp>
#include <Omega_h_timer.hpp>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#ifdef OSHB_REF_OPENMP
#include <omp.h>
#endif

using namespace Omega_h;

struct Dump {
  FILE* f;
  explicit Dump(std::string const& path) {
    f = fopen(path.c_str(), "wb");
    if (!f) {
      fprintf(stderr, "cannot open %s\n", path.c_str());
      exit(1);
    }
    fwrite("OSHD1\n", 1, 6, f);
  }
  ~Dump() { fclose(f); }
  void raw(std::string const& name, int dtype, size_t n, void const* p, size_t esz) {
    unsigned nl = unsigned(name.size());
    fwrite(&nl, 4, 1, f);
    fwrite(name.data(), 1, nl, f);
    unsigned char dt = (unsigned char)dtype;
    fwrite(&dt, 1, 1, f);
    unsigned long long cnt = n;
    fwrite(&cnt, 8, 1, f);
    if (n) fwrite(p, esz, n, f);
  }
  void put(std::string const& name, Read<I8> a) {
    HostRead<I8> h(a);
    raw(name, 0, size_t(h.size()), h.data(), 1);
  }
  void put(std::string const& name, Read<I32> a) {
    HostRead<I32> h(a);
    raw(name, 1, size_t(h.size()), h.data(), 4);
  }
  void put(std::string const& name, Read<I64> a) {
    HostRead<I64> h(a);
    raw(name, 2, size_t(h.size()), h.data(), 8);
  }
  void put(std::string const& name, Read<Real> a) {
    HostRead<Real> h(a);
    raw(name, 3, size_t(h.size()), h.data(), 8);
  }
  void scalar(std::string const& name, long long v) { raw(name, 2, 1, &v, 8); }
  void scalarf(std::string const& name, double v) { raw(name, 3, 1, &v, 8); }
};

static void dump_mesh(Dump& d, std::string const& pre, Mesh* mesh) {
  d.scalar(pre + "dim", mesh->dim());
  for (Int dim = 0; dim <= mesh->dim(); ++dim) {
    auto ds = std::to_string(dim);
    d.scalar(pre + "nents" + ds, mesh->nents(dim));
    if (dim > 0) {
      auto down = mesh->ask_down(dim, dim - 1);
      d.put(pre + "down" + ds, down.ab2b);
      if (down.codes.exists()) d.put(pre + "codes" + ds, down.codes);
    }
    std::string names;
    for (Int i = 0; i < mesh->ntags(dim); ++i) {
      auto tag = mesh->get_tag(dim, i);
      auto key = pre + "tag" + ds + ":" + tag->name();
      d.scalar(key + ":ncomps", tag->ncomps());
      if (is<I8>(tag)) d.put(key, as<I8>(tag)->array());
      if (is<I32>(tag)) d.put(key, as<I32>(tag)->array());
      if (is<I64>(tag)) d.put(key, as<I64>(tag)->array());
      if (is<Real>(tag)) d.put(key, as<Real>(tag)->array());
    }
  }
}

static void dump_derived(Dump& d, std::string const& pre, Mesh* mesh) {
  auto dim = mesh->dim();
  for (Int hd = 2; hd <= dim; ++hd) {
    d.put(pre + "verts_of" + std::to_string(hd), mesh->ask_verts_of(hd));
  }
  if (dim == 3) {
    auto r2e = mesh->ask_down(3, 1);
    d.put(pre + "down31", r2e.ab2b);
    d.put(pre + "codes31", r2e.codes);
  }
  for (Int lo = 0; lo < dim; ++lo) {
    for (Int hi = lo + 1; hi <= dim; ++hi) {
      auto up = mesh->ask_up(lo, hi);
      auto s = std::to_string(lo) + std::to_string(hi);
      d.put(pre + "up" + s + ":a2ab", up.a2ab);
      d.put(pre + "up" + s + ":ab2b", up.ab2b);
      d.put(pre + "up" + s + ":codes", up.codes);
    }
  }
  auto star = mesh->ask_star(EDGE);
  d.put(pre + "star1:a2ab", star.a2ab);
  d.put(pre + "star1:ab2b", star.ab2b);
}

static void set_metric(Mesh* mesh, int n, int kind) {
  auto dim = mesh->dim();
  auto nv = mesh->nverts();
  auto coords = mesh->coords();
  if (kind == 0) {
    Real h = 1.0 / n / 2.0;
    mesh->add_tag(VERT, "metric", 1, Reals(nv, metric_eigenvalue_from_length(h)));
  } else if (kind == 3) {
    Write<Real> m(nv);
    auto f = OMEGA_H_LAMBDA(LO v) {
      Real radius = 0;
      for (Int j = 0; j < dim; ++j) radius += square(coords[v * dim + j]);
      radius = std::sqrt(radius);
      auto coarse = 0.4;
      auto fine = 0.04;
      auto diagonal = std::sqrt(double(3)) - 0.5;
      auto distance = std::abs(radius - 0.5) / diagonal;
      auto h = coarse * distance + fine * (1.0 - distance);
      m[v] = metric_eigenvalue_from_length(h);
    };
    parallel_for(nv, f);
    mesh->add_tag(VERT, "metric", 1, Reals(m));
  } else {
    Write<Real> m(nv * symm_ncomps(dim));
    if (dim == 3) {
      auto f = OMEGA_H_LAMBDA(LO v) {
        auto x = get_vector<3>(coords, v);
        Real t = std::tanh(20.0 * (x[2] - 0.5));
        Real hz = (1.0 / n) * (1.0 - 0.75 * (1.0 - t * t));
        auto hh = vector_3(1.0 / n, (kind == 2 ? 0.7 : 1.0) / n, hz);
        set_symm(m, v, compose_metric(identity_matrix<3, 3>(), hh));
      };
      parallel_for(nv, f);
    } else {
      auto f = OMEGA_H_LAMBDA(LO v) {
        auto x = get_vector<2>(coords, v);
        Real t = std::tanh(20.0 * (x[1] - 0.5));
        Real hy = (1.0 / n) * (1.0 - 0.75 * (1.0 - t * t));
        auto hh = vector_2((kind == 2 ? 0.7 : 1.0) / n, hy);
        /* rotate the frame by 30 degrees so the 2x2 tensors are not diagonal */
        Real c = std::cos(PI / 6.0), s = std::sin(PI / 6.0);
        auto r = matrix_2x2(c, -s, s, c);
        set_symm(m, v, compose_metric(r, hh));
      };
      parallel_for(nv, f);
    }
    mesh->add_tag(VERT, "metric", symm_ncomps(dim), Reals(m));
  }
}

static Mesh make_box(Library* lib, int dim, int n) {
  return build_box(lib->world(), OMEGA_H_SIMPLEX, 1., 1., dim == 3 ? 1. : 0., n, n, dim == 3 ? n : 0);
}

static int mode_refine(Library* lib, int argc, char** argv) {
  int dim = atoi(argv[2]);
  int n = atoi(argv[3]);
  int kind = atoi(argv[4]);
  int npasses = atoi(argv[5]);
  std::string prefix = argv[6];
  auto mesh = make_box(lib, dim, n);
  set_metric(&mesh, n, kind);
  auto opts = AdaptOpts(&mesh);
  opts.verbosity = SILENT;
  if (argc > 7 && atof(argv[7]) > 0) opts.min_quality_allowed = atof(argv[7]);
  bool askq = (argc > 8) ? atoi(argv[8]) != 0 : true;
  bool userfields = (argc > 9) ? atoi(argv[9]) != 0 : false;
  if (userfields) {
    // user fields carried through the pass by TransferOpts::type_map (src/Omega_h_adapt.hpp:30)
    auto coords = mesh.coords();
    auto nv = mesh.nverts();
    Write<Real> temp(nv), auxm(nv);
    auto f = OMEGA_H_LAMBDA(LO v) {
      Real sum = 0;
      for (Int j = 0; j < dim; ++j) sum += (j + 1) * coords[v * dim + j];
      temp[v] = sum;
      auxm[v] = metric_eigenvalue_from_length(0.05 + 0.1 * coords[v * dim]);
    };
    parallel_for(nv, f);
    mesh.add_tag(VERT, "temperature", 1, Reals(temp));
    mesh.add_tag(VERT, "aux_metric", 1, Reals(auxm));
    for (Int d = 0; d <= dim; ++d) {
      auto cid = mesh.get_array<ClassId>(d, "class_id");
      auto cdim = mesh.get_array<I8>(d, "class_dim");
      Write<LO> mat(mesh.nents(d));
      auto g = OMEGA_H_LAMBDA(LO e) { mat[e] = cid[e] * 7 + cdim[e]; };
      parallel_for(mesh.nents(d), g);
      mesh.add_tag(d, "mat_id", 1, LOs(mat));
    }
    Write<Real> rho(mesh.nelems());
    auto h = OMEGA_H_LAMBDA(LO e) { rho[e] = 1.0 + (e % 17) * 0.25; };
    parallel_for(mesh.nelems(), h);
    mesh.add_tag(dim, "rho", 1, Reals(rho));
    opts.xfer_opts.type_map["temperature"] = OMEGA_H_LINEAR_INTERP;
    opts.xfer_opts.type_map["aux_metric"] = OMEGA_H_METRIC;
    opts.xfer_opts.type_map["mat_id"] = OMEGA_H_INHERIT;
    opts.xfer_opts.type_map["rho"] = OMEGA_H_DENSITY;
  }
  for (int pass = 0; npasses < 0 || pass < npasses; ++pass) {
    if (kind == 3 && pass > 0) {
      set_metric(&mesh, n, kind);
    }
    mesh.ask_lengths();
    if (askq) mesh.ask_qualities();
    Dump d(prefix + "_pass" + std::to_string(pass) + ".oshd");
    d.scalarf("opts:max_length_desired", opts.max_length_desired);
    d.scalarf("opts:min_quality_allowed", opts.min_quality_allowed);
    d.scalar("metric_kind", kind);
    d.scalar("box_n", n);
    for (auto const& kv : opts.xfer_opts.type_map) d.scalar("xfer:" + kv.first, int(kv.second));
    dump_mesh(d, "in:", &mesh);
    /* intermediates, recomputed with the reference's own functions on a shallow copy
       (arrays are immutable and shared; src/Omega_h_refine.cpp:17-41) */
    Mesh old = mesh;
    dump_derived(d, "in:", &old);
    auto lengths = old.ask_lengths();
    auto edge_is_cand = each_gt(lengths, opts.max_length_desired);
    d.put("mid:candidate", edge_is_cand);
    bool any = get_max(edge_is_cand) == 1;
    if (any) {
      auto cands2edges = collect_marked(edge_is_cand);
      d.put("mid:cands2edges", cands2edges);
      d.put("mid:mident_metrics",
          get_mident_metrics(&old, EDGE, cands2edges, old.get_array<Real>(VERT, "metric")));
      auto cand_quals = refine_qualities(&old, cands2edges);
      d.put("mid:cand_quals", cand_quals);
      auto good = each_geq_to(cand_quals, opts.min_quality_allowed);
      if (get_max(good) == 1) {
        auto initial = map_onto(good, cands2edges, old.nedges(), I8(0), 1);
        auto edge_quals = map_onto(cand_quals, cands2edges, old.nedges(), 0.0, 1);
        auto keys = find_indset(&old, EDGE, edge_quals, initial);
        d.put("mid:key", keys);
        d.put("mid:rep_vertex2md_order", get_rep2md_order_adapt(&old, EDGE, VERT, keys));
      }
    }
    bool did = refine_by_size(&mesh, opts);
    d.scalar("did", did ? 1 : 0);
    if (did) dump_mesh(d, "out:", &mesh);
    printf("pass %d: did=%d nelems=%d\n", pass, int(did), mesh.nelems());
    if (!did) break;
  }
  return 0;
}

static int mode_time(Library* lib, int argc, char** argv) {
  int dim = atoi(argv[2]);
  int n = atoi(argv[3]);
  int kind = atoi(argv[4]);
  int maxpasses = (argc > 5) ? atoi(argv[5]) : 1000;
  auto mesh = make_box(lib, dim, n);
  set_metric(&mesh, n, kind);
  auto opts = AdaptOpts(&mesh);
  opts.verbosity = SILENT;
  mesh.ask_lengths();
  mesh.ask_qualities();
  int threads = 1;
#ifdef OSHB_REF_OPENMP
  threads = omp_get_max_threads();
#endif
  printf("{\"impl\":\"reference\",\"threads\":%d,\"dim\":%d,\"n\":%d,\"metric\":%d,\"passes\":[", threads, dim, n, kind);
  double total = 0;
  long long first = mesh.nelems();
  for (int pass = 0; pass < maxpasses; ++pass) {
    long long nb = mesh.nelems();
    auto t0 = now();
    bool did = refine_by_size(&mesh, opts);
    auto t1 = now();
    if (!did) break;
    total += (t1 - t0);
    printf("%s{\"before\":%lld,\"after\":%lld,\"seconds\":%.6f}", pass ? "," : "", nb, (long long)mesh.nelems(), t1 - t0);
    fflush(stdout);
  }
  printf("],\"nelems_before\":%lld,\"nelems_after\":%lld,\"seconds\":%.6f}\n", first, (long long)mesh.nelems(), total);
  return 0;
}


// timeloops: the complete `while (refine_by_size)` loop, repeated `reps` times on shallow copies of the same
// input mesh (the reference's Mesh copy shares its immutable arrays), stopping early when `budget_s` seconds
// of wall time are used up. Prints one JSON line per repetition.
static int mode_timeloops(Library* lib, int argc, char** argv) {
  int dim = atoi(argv[2]);
  int n = atoi(argv[3]);
  int kind = atoi(argv[4]);
  int reps = (argc > 5) ? atoi(argv[5]) : 1;
  double budget = (argc > 6) ? atof(argv[6]) : 1e30;
  auto base = make_box(lib, dim, n);
  set_metric(&base, n, kind);
  auto opts = AdaptOpts(&base);
  opts.verbosity = SILENT;
  base.ask_lengths();
  base.ask_qualities();
  int threads = 1;
#ifdef OSHB_REF_OPENMP
  threads = omp_get_max_threads();
#endif
  auto start = now();
  for (int rep = 0; rep < reps; ++rep) {
    Mesh mesh = base;
    long long first = mesh.nelems();
    int passes = 0;
    auto t0 = now();
    while (refine_by_size(&mesh, opts)) ++passes;
    auto t1 = now();
    printf("{\"impl\":\"reference\",\"threads\":%d,\"dim\":%d,\"n\":%d,\"metric\":%d,\"rep\":%d,\"passes\":%d,"
           "\"nelems_before\":%lld,\"nelems_after\":%lld,\"seconds\":%.6f}\n",
        threads, dim, n, kind, rep, passes, first, (long long)mesh.nelems(), t1 - t0);
    fflush(stdout);
    if (now() - start > budget) break;
  }
  return 0;
}

// digest: the complete while(refine_by_size) loop, then one line per array of the final mesh with a
// position-dependent 64-bit checksum (sum over i of (bits_i + 1) * (2 i + 1) mod 2^64 -- order-sensitive, computable
// with wrapping uint64 arithmetic anywhere) plus, for reals, sum / min / max. The full-size parity test compares
// these against the same digests of the GPU result: the reference itself as the oracle at BASELINE's sizes, where
// dumping every array would be gigabytes.
template <typename T>
static unsigned long long digest_bits(HostRead<T> const& a) {
  unsigned long long h = 0;
  for (LO i = 0; i < a.size(); ++i) {
    unsigned long long bits = 0;
    T v = a[i];
    memcpy(&bits, &v, sizeof(T));  // little endian: the low bytes
    h += (bits + 1ull) * (2ull * (unsigned long long)i + 1ull);
  }
  return h;
}
template <typename T>
static void digest_line(std::string const& key, Read<T> arr, int ncomps) {
  HostRead<T> a(arr);
  printf("{\"key\":\"%s\",\"n\":%lld,\"ncomps\":%d,\"hash\":\"%llu\"", key.c_str(), (long long)a.size(), ncomps, digest_bits(a));
  if (std::is_same<T, Real>::value) {
    long double sum = 0;
    double mn = 1e300, mx = -1e300;
    for (LO i = 0; i < a.size(); ++i) {
      double v = double(a[i]);
      sum += v;
      if (v < mn) mn = v;
      if (v > mx) mx = v;
    }
    printf(",\"sum\":%.17g,\"min\":%.17g,\"max\":%.17g", double(sum), mn, mx);
  }
  printf("}\n");
}
static int mode_digest(Library* lib, int argc, char** argv) {
  int dim = atoi(argv[2]);
  int n = atoi(argv[3]);
  int kind = atoi(argv[4]);
  auto mesh = make_box(lib, dim, n);
  set_metric(&mesh, n, kind);
  if (argc > 5) {
    // the input metric, raw doubles, so the other side starts from bit-identical inputs
    HostRead<Real> hm(mesh.get_array<Real>(VERT, "metric"));
    FILE* f = fopen(argv[5], "wb");
    if (!f || fwrite(hm.data(), sizeof(Real), size_t(hm.size()), f) != size_t(hm.size())) return 3;
    fclose(f);
  }
  auto opts = AdaptOpts(&mesh);
  opts.verbosity = SILENT;
  mesh.ask_lengths();
  mesh.ask_qualities();
  int passes = 0;
  while (refine_by_size(&mesh, opts)) ++passes;
  printf("{\"key\":\"passes\",\"n\":%d}\n", passes);
  for (Int d = 0; d <= mesh.dim(); ++d) {
    auto ds = std::to_string(d);
    printf("{\"key\":\"nents%d\",\"n\":%lld}\n", d, (long long)mesh.nents(d));
    if (d > 0) {
      auto down = mesh.ask_down(d, d - 1);
      digest_line<LO>("down" + ds, down.ab2b, d + 1);
      if (down.codes.exists()) digest_line<I8>("codes" + ds, down.codes, d + 1);
    }
    for (Int i = 0; i < mesh.ntags(d); ++i) {
      auto tag = mesh.get_tag(d, i);
      auto key = "tag" + ds + ":" + tag->name();
      if (is<I8>(tag)) digest_line<I8>(key, as<I8>(tag)->array(), tag->ncomps());
      if (is<I32>(tag)) digest_line<I32>(key, as<I32>(tag)->array(), tag->ncomps());
      if (is<I64>(tag)) digest_line<I64>(key, as<I64>(tag)->array(), tag->ncomps());
      if (is<Real>(tag)) digest_line<Real>(key, as<Real>(tag)->array(), tag->ncomps());
    }
  }
  fflush(stdout);
  return 0;
}

// rib: Mesh::balance()'s element -> rank assignment without MPI. recursively_bisect (src/Omega_h_inertia.cpp:162-193)
// marks a group's elements with inertia::mark_bisection, sends unmarked / marked elements to the lower / upper half
// of the group's ranks (bi_partition, src/Omega_h_bipart.cpp:9-33) and recurses on each half; the marks depend on
// the group's elements only through order-independent reductions (repro_sum), so calling the reference's own
// mark_bisection on each group with a one-rank communicator gives the assignment of the MPI run.
static void rib_recurse(CommPtr comm, Reals ecoords, LOs elems, int first, int size, Write<LO> parts,
    std::vector<Vector<3>>* axes) {
  if (size == 1 || elems.size() == 0) {
    auto f = OMEGA_H_LAMBDA(LO i) { parts[elems[i]] = first; };
    parallel_for(elems.size(), f);
    return;
  }
  auto coords = unmap(elems, ecoords, 3);
  Reals masses(elems.size(), 1.0);
  Vector<3> axis;
  auto marks = inertia::mark_bisection(comm, coords, masses, 2.0, axis);
  axes->push_back(axis);
  auto upper = collect_marked(marks);
  auto lower = collect_marked(invert_marks(marks));
  rib_recurse(comm, ecoords, unmap(lower, elems, 1), first, size / 2, parts, axes);
  rib_recurse(comm, ecoords, unmap(upper, elems, 1), first + size / 2, size / 2, parts, axes);
}

static int mode_rib(Library* lib, int, char** argv) {
  int dim = atoi(argv[2]);
  int nx = atoi(argv[3]), ny = atoi(argv[4]), nz = atoi(argv[5]);
  int nparts = atoi(argv[6]);
  auto mesh = build_box(lib->world(), OMEGA_H_SIMPLEX, double(nx), double(ny), dim == 3 ? double(nz) : 0.0, nx, ny,
      dim == 3 ? nz : 0);
  auto ecoords = average_field(&mesh, dim, LOs(mesh.nelems(), 0, 1), dim, mesh.coords());
  if (dim < 3) ecoords = resize_vectors(ecoords, dim, 3);
  Write<LO> parts(mesh.nelems(), -1);
  std::vector<Vector<3>> axes;
  rib_recurse(lib->world(), ecoords, LOs(mesh.nelems(), 0, 1), 0, nparts, parts, &axes);
  Dump d(argv[7]);
  dump_mesh(d, "in:", &mesh);
  d.put("rib:parts", LOs(parts));
  d.scalar("rib:nparts", nparts);
  return 0;
}

// diff: the reference's compare_meshes (src/Omega_h_compare.cpp:179-277) on two .osh files, as src/oshdiff.cpp:78-87 calls it
static int mode_diff(Library* lib, int, char** argv) {
  Mesh a(lib), b(lib);
  binary::read(argv[2], lib->world(), &a);
  binary::read(argv[3], lib->world(), &b);
  auto opts = MeshCompareOpts::init(&a, VarCompareOpts{VarCompareOpts::RELATIVE, atof(argv[4]), atof(argv[5])});
  auto res = compare_meshes(&a, &b, opts, true, true);
  printf("RESULT %d\n", int(res));
  return 0;
}

static int mode_box(Library* lib, int, char** argv) {
  int dim = atoi(argv[2]);
  int n = atoi(argv[3]);
  auto mesh = make_box(lib, dim, n);
  Dump d(argv[4]);
  dump_mesh(d, "in:", &mesh);
  dump_derived(d, "in:", &mesh);
  return 0;
}

static int mode_adjtime(Library* lib, int, char** argv) {
  int n = atoi(argv[2]);
  auto mesh = make_box(lib, 3, n);
  auto rv2v = mesh.ask_verts_of(3);
  auto fv2v = mesh.ask_verts_of(2);
  auto v2f = mesh.ask_up(0, 2);
  auto r2f = mesh.ask_down(3, 2);
  auto t0 = now();
  auto v2r = invert_adj(Adj(rv2v), 4, mesh.nverts(), 3, 0);
  auto t1 = now();
  auto f2r = invert_adj(r2f, 4, mesh.nfaces(), 3, 2);
  auto t2 = now();
  auto down = reflect_down(rv2v, fv2v, v2f, OMEGA_H_SIMPLEX, 3, 2);
  auto t3 = now();
  printf("{\"ntets\":%d,\"invert_adj_r2v_s\":%.6f,\"invert_adj_r2f_s\":%.6f,\"reflect_down_r2f_s\":%.6f}\n",
      mesh.nelems(), t1 - t0, t2 - t1, t3 - t2);
  return 0;
}

// writeosh <dim> <n> <metric> <npasses> <path.osh> <dump>: box + metric, npasses refine passes, then
// binary::write (src/Omega_h_file.cpp:518-540) of the mesh and an OSHD dump of the same mesh
static int mode_writeosh(Library* lib, int, char** argv) {
  int dim = atoi(argv[2]);
  int n = atoi(argv[3]);
  int kind = atoi(argv[4]);
  int npasses = atoi(argv[5]);
  auto mesh = make_box(lib, dim, n);
  set_metric(&mesh, n, kind);
  auto opts = AdaptOpts(&mesh);
  opts.verbosity = SILENT;
  mesh.ask_lengths();
  mesh.ask_qualities();
  for (int pass = 0; pass < npasses; ++pass) {
    if (!refine_by_size(&mesh, opts)) break;
  }
  binary::write(argv[6], &mesh);
  Dump d(argv[7]);
  dump_mesh(d, "in:", &mesh);
  return 0;
}

// readosh <path.osh> <dump>: binary::read (src/Omega_h_file.cpp:574-590) of a mesh somebody else
// wrote, then an OSHD dump of what the reference understood
static int mode_readosh(Library* lib, int, char** argv) {
  Mesh mesh(lib);
  binary::read(argv[2], lib->world(), &mesh);
  Dump d(argv[3]);
  dump_mesh(d, "in:", &mesh);
  return 0;
}

int main(int argc, char** argv) {
  auto lib = Library(&argc, &argv);
  if (argc < 2) {
    fprintf(stderr, "usage: ref_driver refine|time|box|adjtime ...\n");
    return 1;
  }
  std::string mode = argv[1];
  if (mode == "refine") return mode_refine(&lib, argc, argv);
  if (mode == "time") return mode_time(&lib, argc, argv);
  if (mode == "timeloops") return mode_timeloops(&lib, argc, argv);
  if (mode == "digest") return mode_digest(&lib, argc, argv);
  if (mode == "box") return mode_box(&lib, argc, argv);
  if (mode == "diff") return mode_diff(&lib, argc, argv);
  if (mode == "rib") return mode_rib(&lib, argc, argv);
  if (mode == "adjtime") return mode_adjtime(&lib, argc, argv);
  if (mode == "writeosh") return mode_writeosh(&lib, argc, argv);
  if (mode == "readosh") return mode_readosh(&lib, argc, argv);
  fprintf(stderr, "unknown mode %s\n", mode.c_str());
  return 1;
}
