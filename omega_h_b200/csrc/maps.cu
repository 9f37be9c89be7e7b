// Standalone array maps on device pointers: the reference's
// unmap / map_into / expand_into / mark_image / invert_injective_map / compound_maps
// (src/Omega_h_map.cpp:25-37,74-87,104-128,183-214,139-165). Inside the refine pass these are folded
// into their consumers (the gather formulation of rebuild.cu never materialises a mapped copy);
// they exist as entry points for callers that bind the primitive layer.
//
// All are HBM-bound index kernels. One thread per OUTPUT word wherever the output is dense
// (unmap, expand_into, compound_maps): stores are coalesced and written once, the reads are the
// gathers. map_into scatters by definition: one thread per input word, coalesced reads.
// Algorithmic bytes: n*(4 + 2*width*sizeof T) (SURVEY 8d "gather width k of T").
#include "mesh.hpp"

namespace oshb {

template <class T>
static void unmap_t(LO const* a2b, int64_t na, T const* b_data, int width, T* a_out) {
  algo_bytes(na * (4 + 2 * int64_t(width) * int64_t(sizeof(T))));
  if (width == 1) {
    parallel_for(na, OSHB_LAMBDA(LO a) { a_out[a] = b_data[a2b[a]]; }, "unmap");
    return;
  }
  parallel_for(na * width, OSHB_LAMBDA(LO i) {
    LO a = i / width;
    int j = i - a * width;
    a_out[i] = b_data[int64_t(a2b[a]) * width + j];
  }, "unmap");
}

template <class T>
static void map_into_t(T const* a_data, LO const* a2b, int64_t na, T* b_data, int width) {
  algo_bytes(na * (4 + 2 * int64_t(width) * int64_t(sizeof(T))));
  if (width == 1) {
    parallel_for(na, OSHB_LAMBDA(LO a) { b_data[a2b[a]] = a_data[a]; }, "map_into");
    return;
  }
  parallel_for(na * width, OSHB_LAMBDA(LO i) {
    LO a = i / width;
    int j = i - a * width;
    b_data[int64_t(a2b[a]) * width + j] = a_data[i];
  }, "map_into");
}

// expand_into: b_data[b] = a_data[a] for every b in [a2b[a], a2b[a+1]). One thread per b (dense,
// coalesced output); its source a = last offset <= b, found by bisection of the offsets (they stay in
// L2: na+1 words against nb*width outputs). Neighbouring b share a, so the a_data reads broadcast.
template <class T>
static void expand_into_t(T const* a_data, LO const* a2b, int64_t na, int64_t nb, T* b_data, int width) {
  algo_bytes((na + 1) * 4 + (na + nb) * int64_t(width) * int64_t(sizeof(T)));
  parallel_for(nb, OSHB_LAMBDA(LO b) {
    LO lo = 0, hi = LO(na);  // invariant: a2b[lo] <= b < a2b[hi]
    while (hi - lo > 1) {
      LO mid = lo + ((hi - lo) >> 1);
      if (a2b[mid] <= b) lo = mid;
      else hi = mid;
    }
    for (int j = 0; j < width; ++j) b_data[int64_t(b) * width + j] = a_data[int64_t(lo) * width + j];
  }, "expand_into");
}

void unmap_bytes(LO const* a2b, int64_t na, void const* b_data, int width, int elem_bytes, void* a_out) {
  if (elem_bytes == 1) unmap_t<I8>(a2b, na, static_cast<I8 const*>(b_data), width, static_cast<I8*>(a_out));
  else if (elem_bytes == 4) unmap_t<LO>(a2b, na, static_cast<LO const*>(b_data), width, static_cast<LO*>(a_out));
  else if (elem_bytes == 8) unmap_t<GO>(a2b, na, static_cast<GO const*>(b_data), width, static_cast<GO*>(a_out));
  else fail(__FILE__, __LINE__, "unmap: element size must be 1, 4 or 8 bytes");
}
void map_into_bytes(void const* a_data, LO const* a2b, int64_t na, void* b_data, int width, int elem_bytes) {
  if (elem_bytes == 1) map_into_t<I8>(static_cast<I8 const*>(a_data), a2b, na, static_cast<I8*>(b_data), width);
  else if (elem_bytes == 4) map_into_t<LO>(static_cast<LO const*>(a_data), a2b, na, static_cast<LO*>(b_data), width);
  else if (elem_bytes == 8) map_into_t<GO>(static_cast<GO const*>(a_data), a2b, na, static_cast<GO*>(b_data), width);
  else fail(__FILE__, __LINE__, "map_into: element size must be 1, 4 or 8 bytes");
}
void expand_into_bytes(void const* a_data, LO const* a2b, int64_t na, int64_t nb, void* b_data, int width, int elem_bytes) {
  if (elem_bytes == 1) expand_into_t<I8>(static_cast<I8 const*>(a_data), a2b, na, nb, static_cast<I8*>(b_data), width);
  else if (elem_bytes == 4) expand_into_t<LO>(static_cast<LO const*>(a_data), a2b, na, nb, static_cast<LO*>(b_data), width);
  else if (elem_bytes == 8) expand_into_t<GO>(static_cast<GO const*>(a_data), a2b, na, nb, static_cast<GO*>(b_data), width);
  else fail(__FILE__, __LINE__, "expand_into: element size must be 1, 4 or 8 bytes");
}

// mark_image (src/Omega_h_map.cpp:183-190): marks[b] = 1 for every b in the image of a2b
void mark_image(LO const* a2b, int64_t na, int64_t nb, I8* marks) {
  dev_memset(marks, 0, size_t(nb));
  parallel_for(na, OSHB_LAMBDA(LO a) { marks[a2b[a]] = 1; }, "mark_image");
}

// invert_injective_map (src/Omega_h_map.cpp:199-205): b2a[a2b[a]] = a, -1 elsewhere
void invert_injective_map(LO const* a2b, int64_t na, int64_t nb, LO* b2a) {
  fill<LO>(b2a, nb, -1);
  parallel_for(na, OSHB_LAMBDA(LO a) { b2a[a2b[a]] = a; }, "invert_injective_map");
}

// compound_maps (src/Omega_h_map.cpp:157-165): a2c[a] = b2c[a2b[a]]
void compound_maps(LO const* a2b, int64_t na, LO const* b2c, LO* a2c) {
  parallel_for(na, OSHB_LAMBDA(LO a) { a2c[a] = b2c[a2b[a]]; }, "compound_maps");
}

}  // namespace oshb
