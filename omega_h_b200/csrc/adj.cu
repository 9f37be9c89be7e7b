// Adjacency derivation kernels: invert_adj, transit, reflect_down, edge star,
// form_uses / find_unique, plus the scan/compaction maps they need.
// Outputs are bit-identical to the reference's (src/Omega_h_adj.cpp, src/Omega_h_map.cpp)
// but the algorithms are re-designed for HBM: no materialised use lists, no per-use
// linear search through vertex stars.
#include "mesh.hpp"
#include "sortnet.hpp"

namespace oshb {


// ---------------------------------------------------------------------------------------
// device error cell
// ---------------------------------------------------------------------------------------
int* device_error_cell() { return reinterpret_cast<int*>(static_cast<char*>(ctx().dscratch) + 1024); }

void device_error_reset() {
  int z = 0;
  h2d(device_error_cell(), &z, sizeof(int));
}

void device_error_check(char const* where) {
  int v = read_scalar(device_error_cell());
  if (v != 0) {
    device_error_reset();
    fail(__FILE__, __LINE__, std::string("device-side check failed (bits ") + std::to_string(v) + ") at " + where);
  }
}

// ---------------------------------------------------------------------------------------
// maps
// ---------------------------------------------------------------------------------------
LOs offset_scan(Bytes a) {
  LOs out(a.size() + 1);
  scan_offsets(a.data(), a.size(), out.data());
  return out;
}
LOs offset_scan(LOs a) {
  LOs out(a.size() + 1);
  scan_offsets(a.data(), a.size(), out.data());
  return out;
}
LO last_of(LOs a) {
  OSHB_CHECK(a.size() > 0);
  return read_scalar(a.data() + (a.size() - 1));
}

// order-preserving stream compaction: scan of the marks + guarded scatter
LOs collect_marked(Bytes marks, LO* count_out) {
  auto n = marks.size();
  auto offsets = offset_scan(marks);
  LO nmarked = last_of(offsets);
  LOs out(nmarked);
  I8 const* m = marks.data();
  LO const* off = offsets.data();
  LO* o = out.data();
  parallel_for(n, OSHB_LAMBDA(LO i) {
    if (m[i]) o[off[i]] = i;
  }, "collect_marked");
  if (count_out) *count_out = nmarked;
  return out;
}

// ---------------------------------------------------------------------------------------
// invert_adj: upward adjacency from downward (src/Omega_h_adj.cpp:178-263).
//   1. count     : degree of every low (integer reductions, no return value)
//   2. offset scan of the degrees
//   3. place     : slot claim (the degree array is counted back down: no cursor array); the use is
//                  stored as ONE 8-byte word (high index << 8 | upward code) so that the scattered
//                  traffic of the pass is one store per use; rows are complete, in arrival order
//   4. sort rows : a CTA takes a fixed range of rows (about 3072 entries), stages it in shared
//                  memory (coalesced), marks the row heads, and every entry finds its rank inside
//                  its row by counting the smaller high indices of that row (rows are short, 2..~30:
//                  a rank sort has no divergence, no per-thread row buffers, and neighbouring lanes
//                  read the same row = shared-memory broadcasts); the sorted range leaves through a
//                  second staging buffer, coalesced. This is the reference's sort_by_high_index.
// What bounds it on B200 (profiles/): stages 1 and 3 issue one scattered global operation per use
// and lane; the SM retires about one distinct 128-byte line per cycle, i.e. ~0.3 T scattered
// lanes/s for the chip, whatever DRAM could deliver.
// Algorithmic bytes: in 4*N*d (+N*d codes); out 4*(L+1) + 4*N*d + N*d.
// ---------------------------------------------------------------------------------------
static constexpr int IA_SPAN = 3072;  // entries a CTA aims for
static constexpr int IA_CAP = 4608;   // staging capacity; denser row ranges are sorted in global memory

#ifndef OSHB_EMU
__global__ void __launch_bounds__(256) k_invert_sort_rows(LO const* __restrict__ off, LO nlows, int rows_per_cta,
    unsigned long long* __restrict__ packed, LO* __restrict__ h_out, I8* __restrict__ c_out) {
  __shared__ unsigned s_h[IA_CAP];
  __shared__ unsigned s_ho[IA_CAP];
  __shared__ I8 s_c[IA_CAP];
  __shared__ I8 s_co[IA_CAP];
  int const t = threadIdx.x;
  int64_t const r0 = int64_t(blockIdx.x) * rows_per_cta;
  int64_t const r1 = (r0 + rows_per_cta < nlows) ? (r0 + rows_per_cta) : int64_t(nlows);
  int64_t const e0 = off[r0];
  int64_t const e1 = off[r1];
  int const ne = int(e1 - e0);
  if (ne == 0) return;
  if (ne > IA_CAP) {
    // unusually dense rows: one thread per row, insertion sort in place in global memory
    for (int64_t r = r0 + t; r < r1; r += 256) {
      int64_t b = off[r], e = off[r + 1];
      for (int64_t i = b + 1; i < e; ++i) {
        unsigned long long x = packed[i];
        int64_t j = i - 1;
        while (j >= b && packed[j] > x) {
          packed[j + 1] = packed[j];
          --j;
        }
        packed[j + 1] = x;
      }
      for (int64_t i = b; i < e; ++i) {
        h_out[i] = LO(packed[i] >> 8);
        c_out[i] = I8(packed[i] & 0xffu);
      }
    }
    return;
  }
  for (int j = t; j < ne; j += 256) {
    unsigned long long x = packed[e0 + j];
    s_h[j] = unsigned(x >> 8);
    s_c[j] = I8(x & 0xffu);
  }
  __syncthreads();
  for (int64_t r = r0 + t; r < r1; r += 256) {
    LO b = off[r];
    if (off[r + 1] > b) s_h[b - e0] |= 0x80000000u;  // head of a non-empty row
  }
  __syncthreads();
  for (int p = t; p < ne; p += 256) {
    unsigned const hp = s_h[p] & 0x7fffffffu;
    int cnt = 0;
    int q = p;
    while (!(s_h[q] >> 31)) {
      --q;
      cnt += ((s_h[q] & 0x7fffffffu) <= hp) ? 1 : 0;
    }
    int const start = q;
    q = p + 1;
    while (q < ne && !(s_h[q] >> 31)) {
      cnt += ((s_h[q] & 0x7fffffffu) < hp) ? 1 : 0;
      ++q;
    }
    s_ho[start + cnt] = hp;
    s_co[start + cnt] = s_c[p];
  }
  __syncthreads();
  for (int j = t; j < ne; j += 256) {
    h_out[e0 + j] = LO(s_ho[j]);
    c_out[e0 + j] = s_co[j];
  }
}
#endif

Adj invert_adj(Adj const& down, int nlows_per_high, LO nlows) {
  int64_t const nhl = down.ab2b.size();
  LOs degrees = filled<LO>(nlows, 0);
  LO const* hl2l = down.ab2b.data();
  LO* deg = degrees.data();
  algo_bytes(nhl * 4 + int64_t(nlows) * 4);
  parallel_for(nhl, OSHB_LAMBDA(LO hl) { atomic_add(&deg[hl2l[hl]], 1); }, "invert_adj(count)");
  LOs l2lh = offset_scan(degrees);
  LO const* off = l2lh.data();
  DArr<unsigned long long> arrival(nhl);
  unsigned long long* pk = arrival.data();
  I8 const* dcodes = down.codes.exists() ? down.codes.data() : nullptr;
  int const deg_h = nlows_per_high;
  algo_bytes(nhl * (4 + (dcodes ? 1 : 0) + 8));
  parallel_for(nhl, OSHB_LAMBDA(LO hl) {
    LO l = hl2l[hl];
    LO h = hl / deg_h;
    int which_down = hl - h * deg_h;
    LO j = atomic_add(&deg[l], -1);  // counts back down to zero: no second cursor array
    I8 code;
    if (dcodes) {
      I8 dc = dcodes[hl];
      code = make_code(code_is_flipped(dc), code_rotation(dc), which_down);
    } else {
      code = make_code(false, 0, which_down);
    }
    pk[int64_t(off[l]) + j - 1] =
        (static_cast<unsigned long long>(static_cast<unsigned>(h)) << 8) | static_cast<unsigned long long>(static_cast<unsigned char>(code));
  }, "invert_adj(place)");
  degrees.reset();
  LOs lh2h(nhl);
  Bytes codes(nhl);
#ifdef OSHB_EMU
  {
    LO* ho = lh2h.data();
    I8* co = codes.data();
    for (LO l = 0; l < nlows; ++l) {
      LO b = off[l], e = off[l + 1];
      for (LO p = b; p < e; ++p) {
        LO cnt = 0;
        for (LO q = b; q < e; ++q)
          if (pk[q] < pk[p] || (pk[q] == pk[p] && q < p)) ++cnt;
        ho[b + cnt] = LO(pk[p] >> 8);
        co[b + cnt] = I8(pk[p] & 0xffu);
      }
    }
    ctx().launches++;
  }
#else
  if (nhl > 0 && nlows > 0) {
    Ctx& c = ctx();
    int64_t rows = int64_t(IA_SPAN) * nlows / nhl;  // rows whose entries fill one staging span on average
    if (rows < 1) rows = 1;
    if (rows > IA_SPAN) rows = IA_SPAN;
    unsigned const blocks = unsigned((int64_t(nlows) + rows - 1) / rows);
    algo_bytes(nhl * 13 + int64_t(nlows) * 4);
    if (c.prof_on) prof_begin("invert_adj(sort rows)");
    k_invert_sort_rows<<<blocks, 256, 0, c.stream>>>(off, nlows, int(rows), pk, lh2h.data(), codes.data());
    OSHB_CUDA(cudaGetLastError());
    if (c.prof_on) prof_end("invert_adj(sort rows)");
    c.launches++;
  }
#endif
  Adj up;
  up.a2ab = l2lh;
  up.ab2b = lh2h;
  up.codes = codes;
  return up;
}

// ---------------------------------------------------------------------------------------
// transit: two-level downward adjacency through the upward template and the alignment
// algebra (src/Omega_h_adj.cpp:443-510). One thread per high entity.
// ---------------------------------------------------------------------------------------
Adj transit(Adj const& h2m, Adj const& m2l, int high_dim, int low_dim) {
  OSHB_CHECK(low_dim == 0 || low_dim == 1);
  int const mid_dim = low_dim + 1;
  OSHB_CHECK(high_dim > mid_dim);
  int const nmids_per_high = simplex_degree(high_dim, mid_dim);
  int const nlows_per_mid = simplex_degree(mid_dim, low_dim);
  int const nlows_per_high = simplex_degree(high_dim, low_dim);
  int64_t const nhighs = h2m.ab2b.size() / nmids_per_high;
  LOs hl2l(nhighs * nlows_per_high);
  Bytes codes;
  if (low_dim == 1) codes = Bytes(nhighs * nlows_per_high);
  LO const* hm2m = h2m.ab2b.data();
  I8 const* m2hm_codes = h2m.codes.data();
  LO const* ml2l = m2l.ab2b.data();
  I8 const* ml_codes = m2l.codes.exists() ? m2l.codes.data() : nullptr;
  LO* out = hl2l.data();
  I8* cout = codes.exists() ? codes.data() : nullptr;
  OSHB_CHECK(m2hm_codes != nullptr);
  parallel_for(nhighs, OSHB_LAMBDA(LO h) {
    int64_t const hl_begin = int64_t(h) * nlows_per_high;
    int64_t const hm_begin = int64_t(h) * nmids_per_high;
    for (int hl = 0; hl < nlows_per_high; ++hl) {
      TemplateUp ut = simplex_up_template0(high_dim, low_dim, hl);
      LO m = hm2m[hm_begin + ut.up];
      I8 m2hm_code = m2hm_codes[hm_begin + ut.up];
      I8 hm2m_code = invert_alignment(nlows_per_mid, m2hm_code);
      int ml = align_index(nlows_per_mid, low_dim, ut.which_down, hm2m_code);
      int64_t ml_begin = int64_t(m) * nlows_per_mid;
      out[hl_begin + hl] = ml2l[ml_begin + ml];
      if (low_dim == 1) {
        bool region_face_flipped = code_is_flipped(hm2m_code);
        bool face_edge_flipped = (code_rotation(ml_codes[ml_begin + ml]) == 1);
        bool flipped = region_face_flipped ^ face_edge_flipped ^ ut.is_flipped;
        cout[hl_begin + hl] = make_code(false, int(flipped), 0);
      }
    }
  }, "transit");
  Adj a;
  a.ab2b = hl2l;
  a.codes = codes;
  return a;
}

// ---------------------------------------------------------------------------------------
// form_uses (src/Omega_h_adj.cpp:155-176) -- only used by find_unique / tests;
// reflect_down below never materialises the use list.
// ---------------------------------------------------------------------------------------
LOs form_uses(LOs hv2v, int high_dim, int low_dim) {
  int const nvh = high_dim + 1;
  int const nvl = low_dim + 1;
  int const nlh = simplex_degree(high_dim, low_dim);
  int64_t const nhigh = hv2v.size() / nvh;
  LOs uv2v(nhigh * nlh * nvl);
  LO const* in = hv2v.data();
  LO* out = uv2v.data();
  parallel_for(nhigh * nlh, OSHB_LAMBDA(LO u) {
    LO h = u / nlh;
    int w = u - h * nlh;
    for (int uv = 0; uv < nvl; ++uv) {
      out[int64_t(u) * nvl + uv] = in[int64_t(h) * nvh + simplex_down_template(high_dim, low_dim, w, uv)];
    }
  }, "form_uses");
  return uv2v;
}

// ---------------------------------------------------------------------------------------
// reflect_down: downward adjacency (+codes) of highs given only vertex tuples.
//
// The reference materialises every use and linearly searches all lows around the use's
// first vertex (find_matches_deg, src/Omega_h_adj.cpp:357-396; 52 % of its CPU time).
// The result (the unique low with the same vertex set + the alignment code,
// IsMatch<2>/<3> src/Omega_h_adj.cpp:297-330) does not depend on how it is found, so
// here it is a bucket join:
//   build : every low is filed under its SMALLEST vertex, stored with its other
//           vertices in cyclic order -> rows of ~7 (edges) / ~11 (tris) packed entries
//           (8 B / 16 B each), contiguous per vertex
//   probe : one thread per use rotates the use to its smallest vertex, streams that one
//           row and compares; the code follows from the two rotations and the flip.
// Algorithmic bytes: 4*N*(hd+1) + 4*L*(ld+1) in, 5*N*n_l out; the bucket
// table adds one write + ~one read of (8|16)*L.
// ---------------------------------------------------------------------------------------
// probe, one thread per HIGH entity: its nlh uses are grouped by their smallest vertex (the three
// faces of a tet around its smallest vertex share one bucket row), so a tet streams 2 rows instead
// of 4, a tet's six edges 3 rows instead of 6, and every row entry is loaded once per high. Row
// entries are fetched four at a time before they are compared (independent loads in flight).
template <int HD, int LD>
static void reflect_probe(LO const* hv, int64_t nhigh, LO const* off, LO const* tab, LO* out, I8* cout, int* err) {
  constexpr int NVH = HD + 1;
  constexpr int NVL = LD + 1;
  constexpr int NLH = (HD == 3) ? (LD == 1 ? 6 : 4) : 3;
  parallel_for(nhigh, OSHB_LAMBDA(LO h) {
    LO v[NVH];
#pragma unroll
    for (int k = 0; k < NVH; ++k) v[k] = hv[int64_t(h) * NVH + k];
    LO um_[NLH], m_[NLH], ua_[NLH], ub_[NLH], found[NLH];
    I8 code[NLH];
#pragma unroll
    for (int w = 0; w < NLH; ++w) {
      LO uv[NVL];
#pragma unroll
      for (int k = 0; k < NVL; ++k) uv[k] = v[simplex_down_template(HD, LD, w, k)];
      int um = 0;
#pragma unroll
      for (int k = 1; k < NVL; ++k)
        if (uv[k] < uv[um]) um = k;
      um_[w] = um;
      m_[w] = uv[um];
      if (NVL == 2) {
        ua_[w] = uv[1 - um];
        ub_[w] = 0;
      } else {
        ua_[w] = uv[(um + 1) % 3];
        ub_[w] = uv[(um + 2) % 3];
      }
      found[w] = -1;
      code[w] = 0;
    }
#pragma unroll
    for (int w0 = 0; w0 < NLH; ++w0) {
      if (found[w0] >= 0) continue;
      LO const m = m_[w0];
      LO const rb = off[m];
      LO const re = off[m + 1];
      for (LO s = rb; s < re; s += 4) {
        LO e0[4], e1[4], e2[4], e3[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          bool in = s + j < re;
          if (NVL == 2) {
            e0[j] = in ? tab[int64_t(s + j) * 2 + 0] : -1;
            e1[j] = in ? tab[int64_t(s + j) * 2 + 1] : 0;
            e2[j] = 0;
            e3[j] = 0;
          } else {
            int4 q = in ? reinterpret_cast<int4 const*>(tab)[s + j] : make_int4(-1, -1, 0, 0);
            e0[j] = q.x;
            e1[j] = q.y;
            e2[j] = q.z;
            e3[j] = q.w;
          }
        }
        bool all = true;
#pragma unroll
        for (int w = 0; w < NLH; ++w) {
          if (w < w0 || m_[w] != m) continue;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (NVL == 2) {
              if (e0[j] == ua_[w]) {
                found[w] = e1[j] >> 1;
                int jm = e1[j] & 1;
                // which_down = position in the low of the use's first vertex
                int which_down = (um_[w] == 0) ? jm : (1 - jm);
                code[w] = make_code(false, which_down, 0);
              }
            } else {
              bool same = (e0[j] == ua_[w] && e1[j] == ub_[w]);
              bool flip = (e0[j] == ub_[w] && e1[j] == ua_[w]);
              if (same || flip) {
                found[w] = e2[j];
                int jm = e3[j];
                // low's vertex list b: b[jm]=m, b[jm+1]=va, b[jm+2]=vb; position j in b of the use's
                // first vertex: same orientation j = jm - um, flipped j = jm + um
                int um = um_[w];
                int jj = same ? ((jm - um + 3) % 3) : ((jm + um) % 3);
                code[w] = make_code(flip, rotation_to_first(3, jj), 0);
              }
            }
          }
          if (found[w] < 0) all = false;
        }
        if (all) break;
      }
    }
    bool bad = false;
#pragma unroll
    for (int w = 0; w < NLH; ++w) {
      if (found[w] < 0) bad = true;
      out[int64_t(h) * NLH + w] = found[w];
      cout[int64_t(h) * NLH + w] = code[w];
    }
    if (bad) atomic_or_i32(err, 1);
  }, "reflect_down(probe)");
}

#ifndef OSHB_EMU
// The same probe, eight lanes per high entity: the lanes of a group read eight consecutive row
// entries in one request (one or two 128-byte lines per group instead of one line per lane and
// entry -- the LSU retires about one distinct line per cycle, which is what bounds this kernel),
// compare them with the group's pending uses, and share a hit by ballot + shuffle.
template <int HD, int LD>
__global__ void __launch_bounds__(256) k_reflect_probe(LO const* __restrict__ hv, int64_t nhigh,
    LO const* __restrict__ off, LO const* __restrict__ tab, LO* __restrict__ out, I8* __restrict__ cout, int* err) {
  constexpr int NVH = HD + 1;
  constexpr int NVL = LD + 1;
  constexpr int NLH = (HD == 3) ? (LD == 1 ? 6 : 4) : 3;
  int const lane = threadIdx.x & 31;
  int const gl = lane & 7;
  unsigned const gmask = 0xffu << (lane & 24);
  int64_t const ngroups = (int64_t(gridDim.x) * 256) >> 3;
  for (int64_t h = (int64_t(blockIdx.x) * 256 + threadIdx.x) >> 3; h < nhigh; h += ngroups) {
    LO v[NVH];
#pragma unroll
    for (int k = 0; k < NVH; ++k) v[k] = hv[h * NVH + k];
    LO um_[NLH], m_[NLH], ua_[NLH], ub_[NLH], found[NLH];
    I8 code[NLH];
#pragma unroll
    for (int w = 0; w < NLH; ++w) {
      LO uv[NVL];
#pragma unroll
      for (int k = 0; k < NVL; ++k) uv[k] = v[simplex_down_template(HD, LD, w, k)];
      int um = 0;
#pragma unroll
      for (int k = 1; k < NVL; ++k)
        if (uv[k] < uv[um]) um = k;
      um_[w] = um;
      m_[w] = uv[um];
      if (NVL == 2) {
        ua_[w] = uv[1 - um];
        ub_[w] = 0;
      } else {
        ua_[w] = uv[(um + 1) % 3];
        ub_[w] = uv[(um + 2) % 3];
      }
      found[w] = -1;
      code[w] = 0;
    }
#pragma unroll
    for (int w0 = 0; w0 < NLH; ++w0) {
      if (found[w0] >= 0) continue;  // uniform in the group
      LO const m = m_[w0];
      LO const rb = off[m];
      LO const re = off[m + 1];
      for (LO s = rb; s < re; s += 8) {
        bool const in = s + gl < re;
        LO e0 = -1, e1 = -1, e2 = 0, e3 = 0;
        if (in) {
          if (NVL == 2) {
            int2 q = reinterpret_cast<int2 const*>(tab)[s + gl];
            e0 = q.x;
            e1 = q.y;
          } else {
            int4 q = reinterpret_cast<int4 const*>(tab)[s + gl];
            e0 = q.x;
            e1 = q.y;
            e2 = q.z;
            e3 = q.w;
          }
        }
        bool all = true;
#pragma unroll
        for (int w = 0; w < NLH; ++w) {
          if (w < w0 || m_[w] != m || found[w] >= 0) continue;  // uniform in the group
          LO f = -1;
          int c = 0;
          if (NVL == 2) {
            if (in && e0 == ua_[w]) {
              f = e1 >> 1;
              int jm = e1 & 1;
              int which_down = (um_[w] == 0) ? jm : (1 - jm);
              c = make_code(false, which_down, 0);
            }
          } else {
            bool same = in && (e0 == ua_[w] && e1 == ub_[w]);
            bool flip = in && (e0 == ub_[w] && e1 == ua_[w]);
            if (same || flip) {
              f = e2;
              int jm = e3;
              int um = um_[w];
              int jj = same ? ((jm - um + 3) % 3) : ((jm + um) % 3);
              c = make_code(flip, rotation_to_first(3, jj), 0);
            }
          }
          unsigned hit = __ballot_sync(gmask, f >= 0);
          if (hit) {
            int src = __ffs(hit) - 1;
            found[w] = __shfl_sync(gmask, f, src);
            code[w] = I8(__shfl_sync(gmask, c, src));
          } else {
            all = false;
          }
        }
        if (all) break;
      }
    }
    bool bad = false;
#pragma unroll
    for (int w = 0; w < NLH; ++w) {
      if (found[w] < 0) bad = true;
      if (gl == w) {
        out[h * NLH + w] = found[w];
        cout[h * NLH + w] = code[w];
      }
    }
    if (bad && gl == 0) atomic_or_i32(err, 1);
  }
}

template <int HD, int LD>
static void reflect_probe_groups(LO const* hv, int64_t nhigh, LO const* off, LO const* tab, LO* out, I8* cout, int* err) {
  if (nhigh <= 0) return;
  Ctx& c = ctx();
  int64_t blocks = (nhigh * 8 + 255) / 256;
  int64_t cap = int64_t(c.sms) * 32;
  if (blocks > cap) blocks = cap;
  if (c.prof_on) prof_begin("reflect_down(probe)");
  k_reflect_probe<HD, LD><<<unsigned(blocks), 256, 0, c.stream>>>(hv, nhigh, off, tab, out, cout, err);
  OSHB_CUDA(cudaGetLastError());
  if (c.prof_on) prof_end("reflect_down(probe)");
  c.launches++;
}
#endif

Adj reflect_down(LOs hv2v, LOs lv2v, LO nverts, int high_dim, int low_dim) {
  OSHB_CHECK(low_dim == 1 || low_dim == 2);
  OSHB_CHECK(high_dim > low_dim && high_dim <= 3);
  int const nvh = high_dim + 1;
  int const nvl = low_dim + 1;
  int const nlh = simplex_degree(high_dim, low_dim);
  int64_t const nhigh = hv2v.size() / nvh;
  int64_t const nlow = lv2v.size() / nvl;
  OSHB_CHECK(nlow < (int64_t(1) << 30));
  LO const* lv = lv2v.data();
  int* err = device_error_cell();
  // --- build buckets: one atomic per low (its arrival rank in the row is kept, 16 bits)
  LOs degrees = filled<LO>(nverts, 0);
  LO* deg = degrees.data();
  DArr<uint16_t> ranks(nlow);
  uint16_t* rk = ranks.data();
  algo_bytes(nlow * (4 * nvl + 2) + int64_t(nverts) * 4);
  parallel_for(nlow, OSHB_LAMBDA(LO l) {
    LO m = lv[int64_t(l) * nvl];
    for (int j = 1; j < nvl; ++j) {
      LO v = lv[int64_t(l) * nvl + j];
      if (v < m) m = v;
    }
    LO r = atomic_add(&deg[m], 1);
    if (r >= 65535) {
      atomic_or_i32(err, 1);
      r = 65535;
    }
    rk[l] = uint16_t(r);
  }, "reflect_down(count+rank)");
  LOs offsets = offset_scan(degrees);
  degrees.reset();
  LO const* off = offsets.data();
  // entry layout: edges {other, (l<<1)|jm}; tris {va, vb, l, jm}
  int const ewords = (low_dim == 1) ? 2 : 4;
  LOs table(nlow * ewords);
  LO* tab = table.data();
  algo_bytes(nlow * (4 * nvl + 2 + 4 * ewords));
  parallel_for(nlow, OSHB_LAMBDA(LO l) {
    int jm = 0;
    LO m = lv[int64_t(l) * nvl];
    for (int j = 1; j < nvl; ++j) {
      LO v = lv[int64_t(l) * nvl + j];
      if (v < m) {
        m = v;
        jm = j;
      }
    }
    int64_t slot = int64_t(off[m]) + rk[l];
    if (nvl == 2) {
      tab[slot * 2 + 0] = lv[int64_t(l) * 2 + (1 - jm)];
      tab[slot * 2 + 1] = (l << 1) | jm;
    } else {
#ifdef OSHB_EMU
      tab[slot * 4 + 0] = lv[int64_t(l) * 3 + (jm + 1) % 3];
      tab[slot * 4 + 1] = lv[int64_t(l) * 3 + (jm + 2) % 3];
      tab[slot * 4 + 2] = l;
      tab[slot * 4 + 3] = jm;
#else
      reinterpret_cast<int4*>(tab)[slot] =
          make_int4(lv[int64_t(l) * 3 + (jm + 1) % 3], lv[int64_t(l) * 3 + (jm + 2) % 3], l, jm);
#endif
    }
  }, "reflect_down(build)");
  ranks.reset();
  // --- probe
  LOs hl2l(nhigh * nlh);
  Bytes codes(nhigh * nlh);
  algo_bytes(nhigh * (4 * nvh + 5 * nlh) + nlow * 4 * ewords);
#ifdef OSHB_EMU
#define OSHB_PROBE reflect_probe
#else
#define OSHB_PROBE reflect_probe_groups
#endif
  if (high_dim == 3 && low_dim == 2) OSHB_PROBE<3, 2>(hv2v.data(), nhigh, off, tab, hl2l.data(), codes.data(), err);
  else if (high_dim == 3 && low_dim == 1) OSHB_PROBE<3, 1>(hv2v.data(), nhigh, off, tab, hl2l.data(), codes.data(), err);
  else OSHB_PROBE<2, 1>(hv2v.data(), nhigh, off, tab, hl2l.data(), codes.data(), err);
#undef OSHB_PROBE
  Adj a;
  a.ab2b = hl2l;
  a.codes = codes;
  return a;
}

// ---------------------------------------------------------------------------------------
// edge star = edges_across_tris (+) edges_across_tets (src/Omega_h_adj.cpp:532-589,
// add_edges src/Omega_h_graph.cpp:14-40), built in one pass instead of two graphs + merge.
// ---------------------------------------------------------------------------------------
Adj edges_star(int dim, Adj const& f2e, Adj const& e2f, Adj const& r2e, Adj const& e2r) {
  int64_t const ne = e2f.a2ab.size() - 1;
  LOs degrees(ne);
  LO* dg = degrees.data();
  LO const* e2ef = e2f.a2ab.data();
  LO const* e2er = (dim == 3) ? e2r.a2ab.data() : nullptr;
  parallel_for(ne, OSHB_LAMBDA(LO e) {
    LO d = 2 * (e2ef[e + 1] - e2ef[e]);
    if (e2er) d += e2er[e + 1] - e2er[e];
    dg[e] = d;
  }, "edges_star(count)");
  LOs e2ee = offset_scan(degrees);
  degrees.reset();
  LO nee = last_of(e2ee);
  LOs ee2e(nee);
  LO* out = ee2e.data();
  LO const* off = e2ee.data();
  LO const* ef2f = e2f.ab2b.data();
  I8 const* ef_codes = e2f.codes.data();
  LO const* fe2e = f2e.ab2b.data();
  LO const* er2r = (dim == 3) ? e2r.ab2b.data() : nullptr;
  I8 const* er_codes = (dim == 3) ? e2r.codes.data() : nullptr;
  LO const* re2e = (dim == 3) ? r2e.ab2b.data() : nullptr;
  parallel_for(ne, OSHB_LAMBDA(LO e) {
    LO k = off[e];
    for (LO ef = e2ef[e]; ef < e2ef[e + 1]; ++ef) {
      LO f = ef2f[ef];
      int ffe = code_which_down(ef_codes[ef]);
      out[k++] = fe2e[int64_t(f) * 3 + (ffe + 1) % 3];
      out[k++] = fe2e[int64_t(f) * 3 + (ffe + 2) % 3];
    }
    if (e2er) {
      for (LO er = e2er[e]; er < e2er[e + 1]; ++er) {
        LO r = er2r[er];
        int rre = code_which_down(er_codes[er]);
        out[k++] = re2e[int64_t(r) * 6 + simplex_opposite_template(REGION, EDGE, rre)];
      }
    }
  }, "edges_star(fill)");
  Adj g;
  g.a2ab = e2ee;
  g.ab2b = ee2e;
  return g;
}

// ---------------------------------------------------------------------------------------
// find_unique (src/Omega_h_adj.cpp:133-153): uses -> canonical orientation -> stable
// sort_by_keys -> jumps -> compaction -> first use of every run, in sorted order.
// ---------------------------------------------------------------------------------------
LOs find_unique(LOs hv2v, int high_dim, int low_dim) {
  OSHB_CHECK(low_dim == 1 || low_dim == 2);
  int const deg = low_dim + 1;
  LOs uv2v = form_uses(hv2v, high_dim, low_dim);
  int64_t const nu = uv2v.size() / deg;
  LOs canon(nu * deg);
  LO const* in = uv2v.data();
  LO* cn = canon.data();
  // get_codes_to_canonical + align_ev2v (src/Omega_h_adj.cpp:71-109): smallest vertex
  // first, then flip so that the vertex after it is the smaller of the remaining two
  parallel_for(nu, OSHB_LAMBDA(LO u) {
    LO v[3];
    for (int k = 0; k < deg; ++k) v[k] = in[int64_t(u) * deg + k];
    int mj = 0;
    for (int k = 1; k < deg; ++k)
      if (v[k] < v[mj]) mj = k;
    LO t[3];
    for (int k = 0; k < deg; ++k) t[k] = v[(mj + k) % deg];
    if (deg == 3 && t[2] < t[1]) {
      LO s = t[1];
      t[1] = t[2];
      t[2] = s;
    }
    for (int k = 0; k < deg; ++k) cn[int64_t(u) * deg + k] = t[k];
  }, "find_unique(canonicalize)");
  LOs sorted2u(nu);
  sort_by_keys(canon.data(), nu, deg, sorted2u.data());
  Bytes jumps(nu);
  I8* jp = jumps.data();
  LO const* s2u = sorted2u.data();
  parallel_for(nu, OSHB_LAMBDA(LO s) {
    if (s == LO(nu) - 1) {
      jp[s] = 1;
      return;
    }
    LO a = s2u[s], b = s2u[s + 1];
    bool eq = true;
    for (int k = 0; k < deg; ++k)
      if (cn[int64_t(a) * deg + k] != cn[int64_t(b) * deg + k]) eq = false;
    jp[s] = eq ? 0 : 1;
  }, "find_unique(jumps)");
  LOs e2sorted = collect_marked(jumps);
  int64_t const ne = e2sorted.size();
  LOs ev2v(ne * deg);
  LO const* e2s = e2sorted.data();
  LO* out = ev2v.data();
  parallel_for(ne, OSHB_LAMBDA(LO e) {
    LO u = s2u[e2s[e]];
    for (int k = 0; k < deg; ++k) out[int64_t(e) * deg + k] = in[int64_t(u) * deg + k];
  }, "find_unique(unmap)");
  return ev2v;
}

}  // namespace oshb
