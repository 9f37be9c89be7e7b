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
// invert_adj: upward adjacency from downward.
//   1. degree histogram of the lows (integer atomics, no return value)
//   2. offset scan
//   3. slot claim (atomics; arrival order is arbitrary)
//   4. rows sorted by high-low use index (== by high index, src/Omega_h_adj.cpp:178-200) and
//      high index + upward code emitted, one CTA per 256 consecutive lows: their rows are one
//      contiguous segment of the slot array, which is staged in shared memory with coalesced
//      loads; eight lanes share a row and every ENTRY ranks itself inside its row by counting the
//      smaller entries (rows are short: the O(len^2) compares run on broadcast 16-byte
//      shared-memory reads) and writes high + code straight to its sorted position. A segment
//      that does not fit the stage, and meshes whose rows are short on average, go through one
//      thread per row (register sorting networks, sortnet.hpp).
// Algorithmic bytes: in 4*N*d (+N*d codes); out 4*(L+1) + 4*N*d + N*d.
// ---------------------------------------------------------------------------------------
OSHB_HD void emit_up(LO hl, int deg_h, I8 const* dcodes, LO* h_out, I8* c_out, int64_t at) {
  LO h = hl / deg_h;
  int which_down = hl - h * deg_h;
  h_out[at] = h;
  if (dcodes) {
    I8 dc = dcodes[hl];
    c_out[at] = make_code(code_is_flipped(dc), code_rotation(dc), which_down);
  } else {
    c_out[at] = make_code(false, 0, which_down);
  }
}

#ifndef OSHB_EMU
constexpr int IA_T = 256;     // threads per CTA
constexpr int IA_R = 256;     // lows per CTA iteration (with vertex -> tets rows of ~24 the segment fills the stage)
constexpr int IA_CAP = 8192;  // staged entries per CTA iteration
constexpr int IA_G = 8;       // lanes that share a row

template <int DEG>
__global__ void __launch_bounds__(IA_T) k_rows_sort(LO const* off, LO* slots, LO nlows, int deg_rt, I8 const* dcodes,
    LO* h_out, I8* c_out) {
  __shared__ __align__(16) LO s_val[IA_CAP];
  __shared__ LO s_off[IA_R + 1];
  int const deg_h = DEG ? DEG : deg_rt;
  int const tid = threadIdx.x;
  int const grp = tid / IA_G, sub = tid % IA_G;
  LO const nblk = (nlows + IA_R - 1) / IA_R;
  for (LO blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    LO const l0 = blk * IA_R;
    int const nrows = (nlows - l0 < IA_R) ? int(nlows - l0) : IA_R;
    for (int r = tid; r <= nrows; r += IA_T) s_off[r] = off[l0 + r];
    __syncthreads();
    LO const seg_b = s_off[0];
    int const n = int(s_off[nrows] - seg_b);
    if (n <= IA_CAP) {
      for (int i = tid; i < n; i += IA_T) s_val[i] = slots[int64_t(seg_b) + i];
      __syncthreads();
      for (int r = grp; r < nrows; r += IA_T / IA_G) {
        int const b = int(s_off[r] - seg_b), e = int(s_off[r + 1] - seg_b);
        for (int i = b + sub; i < e; i += IA_G) {
          LO const v = s_val[i];
          // rank of the entry inside its row: scalar head up to a 16-byte boundary, 4 entries per
          // shared-memory load, scalar tail
          int rank = 0;
          int j = b;
          for (; j < e && (j & 3); ++j) rank += (s_val[j] < v) ? 1 : 0;
          for (; j + 4 <= e; j += 4) {
            int4 const q = *reinterpret_cast<int4 const*>(s_val + j);
            rank += ((q.x < v) ? 1 : 0) + ((q.y < v) ? 1 : 0) + ((q.z < v) ? 1 : 0) + ((q.w < v) ? 1 : 0);
          }
          for (; j < e; ++j) rank += (s_val[j] < v) ? 1 : 0;
          emit_up(v, deg_h, dcodes, h_out, c_out, int64_t(seg_b) + b + rank);
        }
      }
    } else {
      for (int r = tid; r < nrows; r += IA_T) {
        LO const b = s_off[r], e = s_off[r + 1];
        sort_small_row(slots + b, e - b);
        for (LO i = b; i < e; ++i) emit_up(slots[i], deg_h, dcodes, h_out, c_out, i);
      }
    }
    __syncthreads();
  }
}

template <int DEG>
static void launch_rows_sort(LO const* off, LO* slots, LO nlows, int deg_h, I8 const* dcodes, LO* h_out, I8* c_out) {
  Ctx& c = ctx();
  int64_t blocks = (int64_t(nlows) + IA_R - 1) / IA_R;
  int64_t const cap = int64_t(c.sms) * 6;
  if (blocks > cap) blocks = cap;
  if (blocks <= 0) return;
  if (c.prof_on) prof_begin("invert_adj(sort+separate)");
  k_rows_sort<DEG><<<unsigned(blocks), IA_T, 0, c.stream>>>(off, slots, nlows, deg_h, dcodes, h_out, c_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) fail(__FILE__, __LINE__, std::string("kernel launch: ") + cudaGetErrorString(e));
  if (c.prof_on) prof_end("invert_adj(sort+separate)");
  c.launches++;
}
#endif

Adj invert_adj(Adj const& down, int nlows_per_high, LO nlows) {
  int64_t const nhl = down.ab2b.size();
  LOs degrees = filled<LO>(nlows, 0);
  LO const* hl2l = down.ab2b.data();
  LO* deg = degrees.data();
  parallel_for(nhl, OSHB_LAMBDA(LO hl) { atomic_add(&deg[hl2l[hl]], 1); }, "invert_adj(count)");
  LOs l2lh = offset_scan(degrees);
  LOs lh2hl(nhl);
  LO const* off = l2lh.data();
  LO* slots = lh2hl.data();
  parallel_for(nhl, OSHB_LAMBDA(LO hl) {
    LO l = hl2l[hl];
    LO j = atomic_add(&deg[l], -1);  // counts back down to zero: no second cursor array
    slots[off[l] + j - 1] = hl;
  }, "invert_adj(fill)");
  degrees.reset();
  LOs lh2h(nhl);
  Bytes codes(nhl);
  LO* h_out = lh2h.data();
  I8* c_out = codes.data();
  I8 const* dcodes = down.codes.exists() ? down.codes.data() : nullptr;
  int const deg_h = nlows_per_high;
  // short rows (edge -> faces ~5, face -> tets <= 2): one thread per row, sorting network in registers
  // (measured faster than the staged tile there: 2.8 / 4.0 / 3.9 ms against 4.9 / 6.5 / 6.6 at 100 M tets);
  // long rows (vertex -> tets ~24): the staged tile (7.2 -> 2.7 ms)
  bool per_row = nhl < int64_t(nlows) * 10;
#ifdef OSHB_EMU
  per_row = true;
#endif
  if (per_row) {
    parallel_for(nlows, OSHB_LAMBDA(LO l) {
      LO const b = off[l];
      LO const e = off[l + 1];
      sort_small_row(slots + b, e - b);
      for (LO i = b; i < e; ++i) emit_up(slots[i], deg_h, dcodes, h_out, c_out, i);
    }, "invert_adj(sort+separate)");
  }
#ifndef OSHB_EMU
  else switch (deg_h) {
    case 2: launch_rows_sort<2>(off, slots, nlows, deg_h, dcodes, h_out, c_out); break;
    case 3: launch_rows_sort<3>(off, slots, nlows, deg_h, dcodes, h_out, c_out); break;
    case 4: launch_rows_sort<4>(off, slots, nlows, deg_h, dcodes, h_out, c_out); break;
    case 6: launch_rows_sort<6>(off, slots, nlows, deg_h, dcodes, h_out, c_out); break;
    default: launch_rows_sort<0>(off, slots, nlows, deg_h, dcodes, h_out, c_out); break;
  }
#endif
  Adj up;
  up.a2ab = l2lh;
  up.ab2b = lh2h;
  up.codes = codes;
  return up;
}

// ---------------------------------------------------------------------------------------
// transit: the two-level downward adjacency high -> low through the stored high -> mid and
// mid -> low rows (what src/Omega_h_adj.cpp:443-510 computes; R->E from R->F o F->E with codes,
// F->V from F->E o E->V, R->V from R->E o E->V).
//
// Mid-centric formulation, one thread per high entity, everything sized at compile time:
//   1. the high's mid row and its alignment codes are read once (16-byte / 4-byte vector loads for
//      tet -> faces);
//   2. for every mid the code is turned ONCE into a "position word": 2 bits per canonical slot w
//      giving where the mid stores its w-th low as seen from the high (the inverse alignment applied
//      to w). A face shared by three template slots (tet edges 0..2 all come from face 0) is decoded
//      once, not three times;
//   3. every template slot then is one field extract + one gathered load; for low = edges the output
//      code is the parity of three flips (high->mid flip, stored mid->low direction, template flip).
// Algorithmic bytes: 5*N*n_mid in + 5*N*n_low out (+ the gathered mid rows, shared by neighbours).
// ---------------------------------------------------------------------------------------
template <int HD, int LD>
struct TransitShape {
  static constexpr int MD = LD + 1;
  static constexpr int NM = (HD == 3) ? (MD == 2 ? 4 : 6) : 3;  // mids per high
  static constexpr int NLM = MD + 1;                            // lows per mid (simplex)
  static constexpr int NL = (HD == 3) ? (LD == 1 ? 6 : 4) : 3;  // lows per high
  // first upward use of low slot s inside the high: (mid slot, slot of the low inside that mid, flipped)
  // tet edges {f0:2, f0:1, f0:0, f1:2, f2:2, f3:2} all flipped; tet verts {e0:0, e1:0, e2:0, e5:1};
  // tri verts {e0:0, e1:0, e2:0}   (src/Omega_h_simplex.hpp:142-229, first entry of each list)
  static OSHB_HD int mid_slot(int s) {
    if (HD == 3 && LD == 1) return s < 3 ? 0 : s - 2;
    if (HD == 3 && LD == 0) return s < 3 ? s : 5;
    return s;
  }
  static OSHB_HD int low_slot(int s) {
    if (HD == 3 && LD == 1) return s < 3 ? 2 - s : 2;
    if (HD == 3 && LD == 0) return s < 3 ? 0 : 1;
    return 0;
  }
  static constexpr bool template_flip = (HD == 3 && LD == 1);
};

// where a mid of NLM lows stores canonical slot w, for the code the HIGH holds of that mid:
// 2 bits per w. Inverse alignment: a flip is its own inverse, a pure rotation r inverts to n - r.
template <int NLM, int LD>
OSHB_HD unsigned position_word(I8 code) {
  int const rot = code_rotation(code);
  bool const flip = code_is_flipped(code);
  int const back = flip ? rot : mod_small(NLM - rot, NLM);
  unsigned word = 0;
#pragma unroll
  for (int w = 0; w < NLM; ++w) {
    int r = mod_small(w + back, NLM);
    int p = flip ? (LD == 0 ? mod_small(NLM - r, NLM) : NLM - 1 - r) : r;
    word |= unsigned(p) << (2 * w);
  }
  return word;
}

template <int HD, int LD>
static void transit_rows(LO const* hm2m, I8 const* hm_codes, LO const* ml2l, I8 const* ml_codes, int64_t nhighs,
    LO* out, I8* cout) {
  typedef TransitShape<HD, LD> S;
  algo_bytes(nhighs * 5 * (S::NM + S::NL));
  parallel_for(nhighs, OSHB_LAMBDA(LO h) {
    LO mid[S::NM];
    I8 mcode[S::NM];
#ifndef OSHB_EMU
    if (S::NM == 4) {
      int4 q = reinterpret_cast<int4 const*>(hm2m)[h];
      mid[0] = q.x;
      mid[1] = q.y;
      mid[2] = q.z;
      mid[3] = q.w;
      unsigned c4 = reinterpret_cast<unsigned const*>(hm_codes)[h];
#pragma unroll
      for (int j = 0; j < 4; ++j) mcode[j & (S::NM - 1)] = I8((c4 >> (8 * j)) & 0xffu);
    } else
#endif
    {
#pragma unroll
      for (int j = 0; j < S::NM; ++j) {
        mid[j] = hm2m[int64_t(h) * S::NM + j];
        mcode[j] = hm_codes[int64_t(h) * S::NM + j];
      }
    }
    unsigned pos[S::NM];
#pragma unroll
    for (int j = 0; j < S::NM; ++j) pos[j] = position_word<S::NLM, LD>(mcode[j]);
    LO low[S::NL];
    I8 lcode[S::NL];
#pragma unroll
    for (int s = 0; s < S::NL; ++s) {
      int const j = S::mid_slot(s);
      int const p = int((pos[j] >> (2 * S::low_slot(s))) & 3u);
      int64_t const at = int64_t(mid[j]) * S::NLM + p;
      low[s] = ml2l[at];
      if (LD == 1) {
        // the edge runs backwards in the high iff an odd number of: high->face flip, the face stores
        // the edge reversed (rotation 1 of a 2-vertex entity), the template's own flip
        bool rev = code_is_flipped(mcode[j]) ^ (code_rotation(ml_codes[at]) == 1) ^ S::template_flip;
        lcode[s] = make_code(false, int(rev), 0);
      }
    }
#pragma unroll
    for (int s = 0; s < S::NL; ++s) {
      out[int64_t(h) * S::NL + s] = low[s];
      if (LD == 1) cout[int64_t(h) * S::NL + s] = lcode[s];
    }
  }, "transit");
}

Adj transit(Adj const& h2m, Adj const& m2l, int high_dim, int low_dim) {
  OSHB_CHECK(low_dim == 0 || low_dim == 1);
  OSHB_CHECK(high_dim > low_dim + 1 && high_dim <= 3);
  OSHB_CHECK(h2m.codes.exists());
  int const nmids_per_high = simplex_degree(high_dim, low_dim + 1);
  int const nlows_per_high = simplex_degree(high_dim, low_dim);
  int64_t const nhighs = h2m.ab2b.size() / nmids_per_high;
  Adj a;
  a.ab2b = LOs(nhighs * nlows_per_high);
  if (low_dim == 1) {
    OSHB_CHECK(high_dim == 3 && m2l.codes.exists());
    a.codes = Bytes(nhighs * nlows_per_high);
    transit_rows<3, 1>(h2m.ab2b.data(), h2m.codes.data(), m2l.ab2b.data(), m2l.codes.data(), nhighs, a.ab2b.data(),
        a.codes.data());
  } else if (high_dim == 3) {
    transit_rows<3, 0>(h2m.ab2b.data(), h2m.codes.data(), m2l.ab2b.data(), nullptr, nhighs, a.ab2b.data(), nullptr);
  } else {
    transit_rows<2, 0>(h2m.ab2b.data(), h2m.codes.data(), m2l.ab2b.data(), nullptr, nhighs, a.ab2b.data(), nullptr);
  }
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
// probe: one thread per use rotates the use to its smallest vertex, streams that one bucket row (one vector
// load per entry) and compares; the code follows from the two rotations and the flip. A one-thread-per-high
// variant (one row per distinct smallest vertex: 2 rows per tet instead of 4, rows read once for three uses)
// was built and measured SLOWER (14.1 against 10.3 ms at 100 M tets): the kernel is bound by the number of
// distinct 32-byte sectors a warp asks of L1 per instruction, and the four uses of a tet in adjacent lanes
// share theirs.
template <int high_dim, int low_dim>
static void reflect_probe(LO const* hv, LO const* off, LO const* tab, int64_t nhigh, LO* out, I8* cout, int* err) {
  constexpr int nvh = high_dim + 1;
  constexpr int nvl = low_dim + 1;
  constexpr int nlh = (high_dim == 3) ? (low_dim == 2 ? 4 : 6) : 3;
  parallel_for(nhigh * nlh, OSHB_LAMBDA(LO u) {
    LO h = u / nlh;
    int w = u - h * nlh;
    LO uv[3];
#pragma unroll
    for (int k = 0; k < nvl; ++k) uv[k] = hv[int64_t(h) * nvh + simplex_down_template(high_dim, low_dim, w, k)];
    // position of the use's smallest vertex (selects, no indexed register array)
    int um = 0;
    LO m = uv[0];
#pragma unroll
    for (int k = 1; k < nvl; ++k)
      if (uv[k] < m) {
        m = uv[k];
        um = k;
      }
    LO const rb = off[m];
    LO const re = off[m + 1];
    LO found = -1;
    I8 code = 0;
    if (nvl == 2) {
      LO const other = (um == 0) ? uv[1] : uv[0];
      for (LO s = rb; s < re; ++s) {
#ifdef OSHB_EMU
        LO const qx = tab[int64_t(s) * 2], qy = tab[int64_t(s) * 2 + 1];
#else
        int2 const q = reinterpret_cast<int2 const*>(tab)[s];
        LO const qx = q.x, qy = q.y;
#endif
        if (qx == other) {
          found = qy >> 1;
          int jm = qy & 1;
          // which_down = position in the low of the use's first vertex
          code = make_code(false, (um == 0) ? jm : (1 - jm), 0);
          break;
        }
      }
    } else {
      LO const ua = (um == 0) ? uv[1] : ((um == 1) ? uv[2 % nvl] : uv[0]);
      LO const ub = (um == 0) ? uv[2 % nvl] : ((um == 1) ? uv[0] : uv[1]);
      for (LO s = rb; s < re; ++s) {
#ifdef OSHB_EMU
        LO const qx = tab[int64_t(s) * 4], qy = tab[int64_t(s) * 4 + 1], qz = tab[int64_t(s) * 4 + 2],
                 qw = tab[int64_t(s) * 4 + 3];
#else
        int4 const q = reinterpret_cast<int4 const*>(tab)[s];
        LO const qx = q.x, qy = q.y, qz = q.z, qw = q.w;
#endif
        bool same = (qx == ua && qy == ub);
        bool flip = (qx == ub && qy == ua);
        if (same || flip) {
          found = qz;
          int jm = qw;
          // low's vertex list b: b[jm]=m, b[jm+1]=va, b[jm+2]=vb.
          // position j in b of the use's first vertex uv[0]:
          //   same orientation: uv[0] = uv[um - um] sits um steps before m  -> j = jm - um
          //   flipped         : walking the use forward walks the low backward -> j = jm + um
          int j = same ? mod_small(jm - um + 3, 3) : mod_small(jm + um, 3);
          code = make_code(flip, rotation_to_first(3, j), 0);
          break;
        }
      }
    }
    if (found < 0) atomic_or_i32(err, 1);
    out[u] = found;
    cout[u] = code;
  }, "reflect_down(probe)");
}

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
  if (high_dim == 3 && low_dim == 2) reflect_probe<3, 2>(hv2v.data(), off, tab, nhigh, hl2l.data(), codes.data(), err);
  else if (high_dim == 3 && low_dim == 1) reflect_probe<3, 1>(hv2v.data(), off, tab, nhigh, hl2l.data(), codes.data(), err);
  else reflect_probe<2, 1>(hv2v.data(), off, tab, nhigh, hl2l.data(), codes.data(), err);
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
