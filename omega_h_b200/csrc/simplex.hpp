// Canonical simplex templates and the 7-bit alignment codes, as arithmetic / tiny
// lookups usable inside kernels. Conventions follow the reference exactly:
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
OSHB_HD int rotate_index(int n, int index, int rotation) { return (index + rotation) % n; }
OSHB_HD int invert_rotation(int n, int rotation) { return (n - rotation) % n; }
OSHB_HD int rotation_to_first(int n, int new_first) { return invert_rotation(n, new_first); }
OSHB_HD int flip_vert_index(int n, int i) { return i == 0 ? 0 : (1 + ((n - 1) - 1 - (i - 1))); }
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
  // (d+1 choose k+1)
  if (to_dim == from_dim) return 1;
  if (from_dim == 1) return 2;
  if (from_dim == 2) return 3;
  /* from_dim == 3 */
  return to_dim == 1 ? 6 : 4;
}

// ---- downward templates (parent-local vertex of a boundary simplex) --------------------
// tet edges {01,12,20,03,13,23}; tet faces {0:(0,2,1) 1:(0,1,3) 2:(1,2,3) 3:(2,0,3)};
// tri edges {01,12,20}
OSHB_HD int simplex_down_template(int elem_dim, int bdry_dim, int which_bdry, int which_vert) {
  if (bdry_dim == 0) return which_bdry;
  if (elem_dim == 1) return which_vert; /* edge -> its own 2 verts (bdry_dim 1 of elem 1 unused) */
  if (elem_dim == 2) {
    /* tri edges */
    return (which_bdry + which_vert) % 3;
  }
  if (bdry_dim == 1) {
    /* tet edges packed 2 bits per vertex: e0(0,1) e1(1,2) e2(2,0) e3(0,3) e4(1,3) e5(2,3) */
    unsigned const v0 = 0x240u /*0,1,2 for 0..2; 0,1,2 for 3..5*/;
    (void)v0;
    int a, b;
    if (which_bdry < 3) {
      a = which_bdry;
      b = (which_bdry + 1) % 3;
    } else {
      a = which_bdry - 3;
      b = 3;
    }
    return which_vert == 0 ? a : b;
  }
  /* tet faces */
  switch (which_bdry) {
    case 0:
      return which_vert == 0 ? 0 : (which_vert == 1 ? 2 : 1);
    case 1:
      return which_vert == 0 ? 0 : (which_vert == 1 ? 1 : 3);
    case 2:
      return which_vert == 0 ? 1 : (which_vert == 1 ? 2 : 3);
    default:
      return which_vert == 0 ? 2 : (which_vert == 1 ? 0 : 3);
  }
}

// ---- opposites (src/Omega_h_simplex.hpp:232-311) ----------------------------------------
OSHB_HD int simplex_opposite_template(int elem_dim, int bdry_dim, int which_bdry) {
  if (elem_dim == 3) {
    if (bdry_dim == 0) {
      /* vert -> face {0->2,1->3,2->1,3->0} */
      return which_bdry == 0 ? 2 : (which_bdry == 1 ? 3 : (which_bdry == 2 ? 1 : 0));
    }
    if (bdry_dim == 1) {
      /* edge -> edge {0<->5,1<->3,2<->4} */
      switch (which_bdry) {
        case 0:
          return 5;
        case 1:
          return 3;
        case 2:
          return 4;
        case 3:
          return 1;
        case 4:
          return 2;
        default:
          return 0;
      }
    }
    /* face -> vert {0->3,1->2,2->0,3->1} */
    return which_bdry == 0 ? 3 : (which_bdry == 1 ? 2 : (which_bdry == 2 ? 0 : 1));
  }
  if (elem_dim == 2) {
    if (bdry_dim == 0) return (which_bdry + 1) % 3; /* vert -> edge {0->1,1->2,2->0} */
    return (which_bdry + 2) % 3;                      /* edge -> vert {0->2,1->0,2->1} */
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
    if (which_bdry < 3) {
      t.up = 0;
      t.which_down = 2 - which_bdry;
    } else {
      t.up = which_bdry - 2;
      t.which_down = 2;
    }
    t.is_flipped = true;
    return t;
  }
  if (elem_dim == 3 && bdry_dim == 0) {
    /* v0:(0,0,0) v1:(1,0,0) v2:(2,0,0) v3:(5,1,0) */
    if (which_bdry < 3) {
      t.up = which_bdry;
      t.which_down = 0;
    } else {
      t.up = 5;
      t.which_down = 1;
    }
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
