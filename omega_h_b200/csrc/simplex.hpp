// Canonical simplex templates and the 7-bit alignment codes, as arithmetic on packed
// literals (2-3 instructions per lookup, no branches, no constant-memory serialisation when
// the index differs per lane). Conventions follow the reference exactly:
//   templates            src/Omega_h_simplex.hpp:23-311
//   alignment codes      src/Omega_h_align.hpp:41-134
#pragma once
#include "rt.hpp"

namespace oshb {

enum { VERT = 0, EDGE = 1, FACE = 2, REGION = 3 };

// ---- alignment codes: (which_down << 3) | (rotation << 1) | flip ----------------------
OSHB_HD bool code_is_flipped(I8 code) { return code & 1; }
OSHB_HD int code_rotation(I8 code) { return (code >> 1) & 3; }
OSHB_HD int code_which_down(I8 code) { return (code >> 3); }
OSHB_HD I8 make_code(bool is_flipped, int rotation, int which_down) {
  return I8((which_down << 3) | (rotation << 1) | int(is_flipped));
}
OSHB_HD int mod_small(int x, int n) { return (x >= n) ? x - n : x; }  // x in [0, 2n)
OSHB_HD int rotate_index(int n, int index, int rotation) { return mod_small(index + rotation, n); }
OSHB_HD int invert_rotation(int n, int rotation) { return mod_small(n - rotation, n); }
OSHB_HD int rotation_to_first(int n, int new_first) { return invert_rotation(n, new_first); }
OSHB_HD int flip_vert_index(int n, int i) { return mod_small(n - i, n); }  // 0 -> 0, i -> n - i
OSHB_HD int flip_edge_index(int n, int i) { return n - 1 - i; }
OSHB_HD int align_vert_index(int n, int index, I8 code) {
  int r = rotate_index(n, index, code_rotation(code));
  return code_is_flipped(code) ? flip_vert_index(n, r) : r;
}
OSHB_HD int align_edge_index(int n, int index, I8 code) {
  int r = rotate_index(n, index, code_rotation(code));
  return code_is_flipped(code) ? flip_edge_index(n, r) : r;
}
OSHB_HD int align_index(int n, int index_dim, int index, I8 code) {
  return index_dim == 0 ? align_vert_index(n, index, code) : align_edge_index(n, index, code);
}
OSHB_HD I8 invert_alignment(int n, I8 code) {
  return code_is_flipped(code) ? code : make_code(false, invert_rotation(n, code_rotation(code)), 0);
}

// ---- degrees --------------------------------------------------------------------------
OSHB_HD int simplex_degree(int from_dim, int to_dim) {
  if (to_dim == from_dim) return 1;
  if (from_dim == 1) return 2;
  if (from_dim == 2) return 3;
  return to_dim == 1 ? 6 : 4;
}

// ---- downward templates (parent-local vertex of a boundary simplex) --------------------
// tet edges {01,12,20,03,13,23}; tet faces {0:(0,2,1) 1:(0,1,3) 2:(1,2,3) 3:(2,0,3)};
// tri edges {01,12,20}
OSHB_HD int simplex_down_template(int elem_dim, int bdry_dim, int which_bdry, int which_vert) {
  if (bdry_dim == 0) return which_bdry;
  if (elem_dim == 1) return which_vert;
  if (elem_dim == 2) return mod_small(which_bdry + which_vert, 3);
  if (bdry_dim == 1) {
    // nibble e of 0x210210 = first vertex of tet edge e, of 0x333021 = second vertex
    unsigned tab = which_vert ? 0x333021u : 0x210210u;
    return int((tab >> (4 * which_bdry)) & 0xfu);
  }
  // 2-bit field (3*face + vert) of 0xcb9d18
  return int((0xcb9d18u >> (2 * (3 * which_bdry + which_vert))) & 3u);
}

// ---- opposites (src/Omega_h_simplex.hpp:232-311) ----------------------------------------
OSHB_HD int simplex_opposite_template(int elem_dim, int bdry_dim, int which_bdry) {
  if (elem_dim == 3) {
    // vert -> face {2,3,1,0}; edge -> edge {5,3,4,1,2,0}; face -> vert {3,2,0,1}
    unsigned tab = (bdry_dim == 0) ? 0x0132u : ((bdry_dim == 1) ? 0x021435u : 0x1023u);
    return int((tab >> (4 * which_bdry)) & 0xfu);
  }
  if (elem_dim == 2) {
    // vert -> edge {1,2,0}; edge -> vert {2,0,1}
    return mod_small(which_bdry + (bdry_dim == 0 ? 1 : 2), 3);
  }
  return 1 - which_bdry;
}

// ---- upward template, first (k=0) entry only: (up, which_down, is_flipped) --------------
// src/Omega_h_simplex.hpp:142-229. Used by transit (src/Omega_h_adj.cpp:476).
struct TemplateUp {
  int up;
  int which_down;
  bool is_flipped;
};
OSHB_HD TemplateUp simplex_up_template0(int elem_dim, int bdry_dim, int which_bdry) {
  TemplateUp t;
  if (elem_dim == 3 && bdry_dim == 1) {
    /* e0:(0,2,1) e1:(0,1,1) e2:(0,0,1) e3:(1,2,1) e4:(2,2,1) e5:(3,2,1) */
    t.up = (which_bdry < 3) ? 0 : which_bdry - 2;
    t.which_down = (which_bdry < 3) ? 2 - which_bdry : 2;
    t.is_flipped = true;
    return t;
  }
  if (elem_dim == 3 && bdry_dim == 0) {
    /* v0:(0,0,0) v1:(1,0,0) v2:(2,0,0) v3:(5,1,0) */
    t.up = (which_bdry < 3) ? which_bdry : 5;
    t.which_down = (which_bdry < 3) ? 0 : 1;
    t.is_flipped = false;
    return t;
  }
  /* elem_dim == 2, bdry_dim == 0: v0:(0,0,0) v1:(1,0,0) v2:(2,0,0) */
  t.up = which_bdry;
  t.which_down = 0;
  t.is_flipped = false;
  return t;
}

}  // namespace oshb
