// Host-side Mesh bookkeeping: tag table, adjacency cache, lazy derivation.
// Mirrors src/Omega_h_mesh.cpp (add_tag :132-177, derive_adj :307-345, ask_adj :347-357,
// ask_lengths/ask_qualities :374-388); all heavy lifting is in adj.cu / geom.cu kernels.
#include "mesh.hpp"

#include <cmath>

namespace oshb {

AdaptOpts::AdaptOpts(int dim) {
  // src/Omega_h_adapt.cpp:52-85
  min_length_desired = 1.0 / sqrt(2.0);
  max_length_desired = sqrt(2.0);
  max_length_allowed = max_length_desired * 2.0;
  if (dim == 3) {
    min_quality_allowed = 0.20;
    min_quality_desired = 0.30;
  } else if (dim == 2) {
    min_quality_allowed = 0.30;
    min_quality_desired = 0.40;
  } else {
    min_quality_allowed = 0.0;
    min_quality_desired = 0.0;
  }
  verbosity = 0;
}

Mesh::Mesh() {
  for (int i = 0; i < 4; ++i) {
    has_star_[i] = false;
    for (int j = 0; j < 4; ++j) has_adj_[i][j] = false;
  }
}

void Mesh::set_ents(int ent_dim, Adj const& down) {
  OSHB_CHECK(ent_dim >= 1 && ent_dim <= 3);
  // the reference asserts !has_ents(ent_dim) (src/Omega_h_mesh.cpp:88); re-setting entities here
  // would leave derived adjacencies, the star and tags of this dimension stale
  OSHB_CHECK(!has_adj_[ent_dim][ent_dim - 1]);
  int deg = simplex_degree(ent_dim, ent_dim - 1);
  nents_[ent_dim] = LO(down.ab2b.size() / deg);
  add_adj(ent_dim, ent_dim - 1, down);
}

Adj Mesh::derive_adj(int from, int to) {
  if (from < to) {
    Adj down = ask_adj(to, from);
    return invert_adj(down, simplex_degree(to, from), nents(from));
  } else if (to < from) {
    OSHB_CHECK(to + 1 < from);
    Adj h2m = ask_adj(from, to + 1);
    Adj m2l = ask_adj(to + 1, to);
    return transit(h2m, m2l, from, to);
  }
  fail(__FILE__, __LINE__, "derive_adj: same-dimension adjacency requested through ask_adj");
}

Adj Mesh::ask_adj(int from, int to) {
  OSHB_CHECK(from >= 0 && from <= dim_ && to >= 0 && to <= dim_ && from != to);
  if (has_adj_[from][to]) return adjs_[from][to];
  Adj a = derive_adj(from, to);
  add_adj(from, to, a);
  return a;
}

Adj Mesh::ask_star(int d) {
  OSHB_CHECK(d == EDGE && dim_ >= 2);
  if (has_star_[d]) return star_[d];
  Adj r2e, e2r;
  if (dim_ == 3) {
    r2e = ask_adj(REGION, EDGE);
    e2r = ask_adj(EDGE, REGION);
  }
  Adj g = edges_star(dim_, ask_adj(FACE, EDGE), ask_adj(EDGE, FACE), r2e, e2r);
  star_[d] = g;
  has_star_[d] = true;
  return g;
}

Tag* Mesh::find_tag(int d, std::string const& name) {
  for (auto& t : tags_[d])
    if (t.name == name) return &t;
  return nullptr;
}
Tag const* Mesh::find_tag(int d, std::string const& name) const {
  for (auto const& t : tags_[d])
    if (t.name == name) return &t;
  return nullptr;
}

void Mesh::remove_tag(int d, std::string const& name) {
  for (size_t i = 0; i < tags_[d].size(); ++i) {
    if (tags_[d][i].name == name) {
      tags_[d].erase(tags_[d].begin() + long(i));
      return;
    }
  }
}

void Mesh::add_tag(int d, Tag const& t, bool internal) {
  OSHB_CHECK(d >= 0 && d <= dim_);
  OSHB_CHECK(t.nvalues() == int64_t(nents(d)) * t.ncomps);
  Tag* old = find_tag(d, t.name);
  if (old)
    *old = t;
  else
    tags_[d].push_back(t);
  if (t.name == "global") globals_state_[d] = 0;
  // user-visible changes of coordinates / metric invalidate the cached measures
  // (react_to_set_tag, src/Omega_h_mesh.cpp:167-177)
  if (!internal && d == VERT && (t.name == "coordinates" || t.name == "metric")) {
    remove_tag(EDGE, "length");
    remove_tag(dim_, "quality");
    if (t.name == "coordinates") remove_tag(dim_, "size");
  }
}

void Mesh::add_tag(int d, std::string const& name, int ncomps, Bytes a, bool internal) {
  Tag t;
  t.name = name;
  t.type = TAG_I8;
  t.ncomps = ncomps;
  t.i8 = a;
  add_tag(d, t, internal);
}
void Mesh::add_tag(int d, std::string const& name, int ncomps, LOs a, bool internal) {
  Tag t;
  t.name = name;
  t.type = TAG_I32;
  t.ncomps = ncomps;
  t.i32 = a;
  add_tag(d, t, internal);
}
void Mesh::add_tag(int d, std::string const& name, int ncomps, GOs a, bool internal) {
  Tag t;
  t.name = name;
  t.type = TAG_I64;
  t.ncomps = ncomps;
  t.i64 = a;
  add_tag(d, t, internal);
}
void Mesh::add_tag(int d, std::string const& name, int ncomps, Reals a, bool internal) {
  Tag t;
  t.name = name;
  t.type = TAG_F64;
  t.ncomps = ncomps;
  t.f64 = a;
  add_tag(d, t, internal);
}

static Tag const* need_tag(Mesh const* m, int d, std::string const& name, int type) {
  Tag const* t = m->find_tag(d, name);
  if (!t) fail(__FILE__, __LINE__, "mesh has no tag \"" + name + "\" on dimension " + std::to_string(d));
  if (t->type != type) fail(__FILE__, __LINE__, "tag \"" + name + "\" has a different type");
  return t;
}
Reals Mesh::get_reals(int d, std::string const& name) const { return need_tag(this, d, name, TAG_F64)->f64; }
Bytes Mesh::get_bytes(int d, std::string const& name) const { return need_tag(this, d, name, TAG_I8)->i8; }
LOs Mesh::get_los(int d, std::string const& name) const { return need_tag(this, d, name, TAG_I32)->i32; }
GOs Mesh::get_gos(int d, std::string const& name) const { return need_tag(this, d, name, TAG_I64)->i64; }

int Mesh::metric_ncomps() const { return need_tag(this, VERT, "metric", TAG_F64)->ncomps; }
int Mesh::metric_dim() const {
  int nc = metric_ncomps();
  for (int i = 1; i <= 3; ++i)
    if (nc == (i * (i + 1)) / 2) return i;
  fail(__FILE__, __LINE__, "metric tag has an unsupported number of components");
}

Reals Mesh::ask_lengths() {
  if (!has_tag(EDGE, "length")) {
    Reals lengths = measure_edges_metric(this, LOs(), get_reals(VERT, "metric"));
    add_tag(EDGE, "length", 1, lengths, true);
  }
  return get_reals(EDGE, "length");
}

Reals Mesh::ask_qualities() {
  if (!has_tag(dim_, "quality")) {
    Reals q = measure_qualities(this, LOs(), get_reals(VERT, "metric"));
    add_tag(dim_, "quality", 1, q, true);
  }
  return get_reals(dim_, "quality");
}

bool Mesh::globals_are_identity(int d) {
  if (globals_state_[d] == 0) {
    GOs g = globals(d);
    GO const* gp = g.data();
    int* cell = reinterpret_cast<int*>(static_cast<char*>(ctx().dscratch) + 1216);
    int z = 0;
    h2d(cell, &z, sizeof(int));
    parallel_for_any(nents(d), OSHB_LAMBDA(LO i)->bool { return gp[i] != GO(i); }, cell, 1, "globals_identity");
    globals_state_[d] = (read_scalar(cell) == 0) ? 1 : 2;
  }
  return globals_state_[d] == 1;
}

Mesh Mesh::copy_meta() const {
  Mesh m;
  m.dim_ = dim_;
  m.xfer_rules_ = xfer_rules_;
  return m;
}

UserTransferHook& user_transfer_hook() {
  static UserTransferHook h;
  return h;
}

}  // namespace oshb
