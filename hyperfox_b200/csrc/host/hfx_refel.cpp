// Host-side reference-element table builder (see hfx_refel.h).
// Behavioural spec: reference src/element/ReferenceElement.cpp (node sets :614-1150, face nodes :136-202, modes :231-445,
// Vandermonde :447-455, interpolation :542-577) and src/element/Cubature.cpp (rule lookup :52-59).
#include "hfx_refel.h"

#include <cmath>
#include <map>
#include <mutex>
#include <stdexcept>
#include <tuple>

#include "hfx_tables_data.inc"

namespace hfx {
namespace {

std::mutex g_mutex;  // the reference's lazily initialised static databases are not thread safe; ours are

int max_order(int dim, Geometry g) {
  static const int s[4] = {10, 10, 10, 5}, o[4] = {5, 5, 5, 2};
  return g == kSimplex ? s[dim] : o[dim];
}

// P_n^{(a,b)}(x) by the three-term recurrence.
double jacobiP(int n, double a, double b, double x) {
  if (n == 0) return 1.0;
  double p0 = 1.0, p1 = 0.5 * (a - b + (a + b + 2.0) * x);
  for (int k = 2; k <= n; k++) {
    double kk = k, c = 2.0 * kk + a + b;
    double a1 = 2.0 * kk * (kk + a + b) * (c - 2.0);
    double a2 = (c - 1.0) * (a * a - b * b);
    double a3 = (c - 2.0) * (c - 1.0) * c;
    double a4 = 2.0 * (kk + a - 1.0) * (kk + b - 1.0) * c;
    double p2 = ((a2 + a3 * x) * p1 - a4 * p0) / a1;
    p0 = p1; p1 = p2;
  }
  return p1;
}
double jacobiDP(int n, double a, double b, double x) {
  return n == 0 ? 0.0 : 0.5 * (n + a + b + 1.0) * jacobiP(n - 1, a + 1.0, b + 1.0, x);
}
double ipow(double x, int n) { double r = 1.0; for (int i = 0; i < n; i++) r *= x; return r; }

// dense inverse (Gauss-Jordan, partial pivoting), row-major n x n
std::vector<double> invert(std::vector<double> A, int n) {
  std::vector<double> I((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) I[(size_t)i * n + i] = 1.0;
  for (int k = 0; k < n; k++) {
    int p = k; double mx = std::fabs(A[(size_t)k * n + k]);
    for (int i = k + 1; i < n; i++) if (std::fabs(A[(size_t)i * n + k]) > mx) { mx = std::fabs(A[(size_t)i * n + k]); p = i; }
    if (mx == 0.0) throw std::runtime_error("ReferenceElement : computeInverseVandermonde : singular matrix");
    if (p != k) for (int j = 0; j < n; j++) { std::swap(A[(size_t)k * n + j], A[(size_t)p * n + j]); std::swap(I[(size_t)k * n + j], I[(size_t)p * n + j]); }
    double d = 1.0 / A[(size_t)k * n + k];
    for (int j = 0; j < n; j++) { A[(size_t)k * n + j] *= d; I[(size_t)k * n + j] *= d; }
    for (int i = 0; i < n; i++) {
      if (i == k) continue;
      double f = A[(size_t)i * n + k];
      if (f == 0.0) continue;
      for (int j = 0; j < n; j++) { A[(size_t)i * n + j] -= f * A[(size_t)k * n + j]; I[(size_t)i * n + j] -= f * I[(size_t)k * n + j]; }
    }
  }
  return I;
}

struct Topo {  // principal vertices and sub-entity lists (edges first, then faces), reference ordering
  std::vector<std::vector<double>> v;
  std::vector<std::vector<int>> ents;
};
const Topo& topo(int dim, Geometry g) {
  static std::map<std::pair<int, int>, Topo> db;
  auto key = std::make_pair(dim, (int)g);
  auto it = db.find(key);
  if (it != db.end()) return it->second;
  Topo t;
  if (dim == 1) { t.v = {{-1}, {1}}; t.ents = {{0}, {1}}; }
  else if (g == kSimplex && dim == 2) { t.v = {{-1, -1}, {1, -1}, {-1, 1}}; t.ents = {{0, 1}, {1, 2}, {2, 0}}; }
  else if (g == kSimplex) {
    t.v = {{-1, -1, -1}, {1, -1, -1}, {-1, 1, -1}, {-1, -1, 1}};
    t.ents = {{0, 1}, {1, 2}, {2, 0}, {3, 0}, {3, 1}, {3, 2}, {3, 1, 0}, {2, 1, 3}, {2, 3, 0}, {0, 1, 2}};
  } else if (dim == 2) { t.v = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}}; t.ents = {{0, 1}, {1, 2}, {2, 3}, {3, 0}}; }
  else {
    t.v = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
    t.ents = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {0, 4}, {1, 5}, {2, 6}, {3, 7}, {4, 5}, {5, 6}, {6, 7}, {7, 4},
              {1, 0, 4, 5}, {2, 1, 5, 6}, {3, 2, 6, 7}, {0, 3, 7, 4}, {0, 1, 2, 3}, {5, 4, 7, 6}};
  }
  return db[key] = t;
}

// affine embedding of the q-dim reference entity onto sub-entity `ent` of the dim-dim element: x = T (xi - xi0) + v0
struct Embed {
  int dim, q;
  std::vector<double> T, v0, xi0;
  Embed(int dim_, int q_, Geometry g, const std::vector<int>& ent) : dim(dim_), q(q_) {
    const Topo& lo = topo(q, g);
    const Topo& hi = topo(dim, g);
    std::vector<double> B((size_t)q * q);  // columns: lower-dim vertex l+1 minus vertex 0
    for (int k = 0; k < q; k++) for (int l = 0; l < q; l++) B[(size_t)k * q + l] = lo.v[l + 1][k] - lo.v[0][k];
    std::vector<double> Bi = invert(B, q);
    std::vector<double> E((size_t)dim * q);
    for (int d = 0; d < dim; d++) for (int l = 0; l < q; l++) E[(size_t)d * q + l] = hi.v[ent[l + 1]][d] - hi.v[ent[0]][d];
    T.assign((size_t)dim * q, 0.0);
    for (int d = 0; d < dim; d++) for (int l = 0; l < q; l++) { double s = 0; for (int k = 0; k < q; k++) s += E[(size_t)d * q + k] * Bi[(size_t)k * q + l]; T[(size_t)d * q + l] = s; }
    v0 = hi.v[ent[0]];
    xi0 = lo.v[0];
  }
  void map(const double* xi, const double* origin, double* x) const {
    for (int d = 0; d < dim; d++) { double s = 0; for (int l = 0; l < q; l++) s += T[(size_t)d * q + l] * (xi[l] - origin[l]); x[d] = s + v0[d]; }
  }
};

std::vector<std::vector<int>> index_tuples(int lo, int hi, int size) {  // lexicographic, first index slowest
  std::vector<std::vector<int>> out;
  std::vector<int> cur(size, lo);
  if (hi < lo) return out;
  while (true) {
    out.push_back(cur);
    int k = size - 1;
    while (k >= 0 && cur[k] == hi) { cur[k] = lo; k--; }
    if (k < 0) break;
    cur[k]++;
  }
  return out;
}

const double* lobatto(int order) {
  for (const auto& e : kHfxLobatto) if (e.order == order) return kHfxLobattoData + e.offset;
  throw std::runtime_error("ReferenceElement : determineNodes : no 1-D Lobatto set for this order");
}

std::map<std::tuple<int, int, int>, std::vector<double>> g_nodes;

const std::vector<double>& node_set_locked(int dim, int order, Geometry g) {
  auto key = std::make_tuple(dim, order, (int)g);
  auto it = g_nodes.find(key);
  if (it != g_nodes.end()) return it->second;
  std::vector<double> pts;
  if (dim == 0) {
  } else if (dim == 1) {
    const double* l = lobatto(order);
    pts.assign(l, l + order + 1);
  } else if (order == 0) {
    pts.assign(dim, 0.0);
  } else {
    const Topo& tp = topo(dim, g);
    for (auto& v : tp.v) pts.insert(pts.end(), v.begin(), v.end());
    std::vector<double> x(dim);
    for (int q = 1; q < dim; q++) {
      const std::vector<double>& sub = node_set_locked(q, order, g);
      int nsub = (int)sub.size() / q;
      for (auto& ent : tp.ents) {
        if ((int)ent.size() != (int)topo(q, g).v.size()) continue;
        Embed em(dim, q, g, ent);
        const double* origin = (g == kSimplex) ? em.xi0.data() : sub.data();
        for (int l = 0; l < nsub; l++) { em.map(&sub[(size_t)l * q], origin, x.data()); pts.insert(pts.end(), x.begin(), x.end()); }
      }
    }
    if (order > 1) {
      const double* l0 = lobatto(order);
      std::vector<double> lob(order + 1);  // {-1, interior ascending ..., +1}
      lob[0] = l0[0];
      for (int k = 2; k <= order; k++) lob[k - 1] = l0[k];
      lob[order] = l0[1];
      auto combos = index_tuples(1, order - 1, dim);
      if (g == kOrthotope) {
        for (auto& c : combos) for (int d = 0; d < dim; d++) pts.push_back(lob[c[d]]);
      } else {
        auto coef = [&](const std::vector<int>& c) {
          int sum = 0; for (int v : c) sum += v;
          double x0 = 2.0 + dim * lob[c[0]];
          for (int d = 1; d < dim; d++) x0 -= lob[c[d]];
          x0 -= lob[order - sum];
          x0 *= 1.0 / (dim + 1.0);
          x0 -= 1.0;
          return x0;
        };
        for (auto& c : combos) {
          int sum = 0; for (int v : c) sum += v;
          if (sum >= order) continue;
          if (dim == 2) { pts.push_back(coef({c[0], c[1]})); pts.push_back(coef({c[1], c[0]})); }
          else { pts.push_back(coef({c[0], c[1], c[2]})); pts.push_back(coef({c[1], c[0], c[2]})); pts.push_back(coef({c[2], c[1], c[0]})); }
        }
      }
    }
    // drop later duplicates (tolerance 1e-12 per coordinate), keeping first occurrences
    int n = (int)pts.size() / dim;
    std::vector<char> dup(n, 0);
    for (int i = 0; i < n; i++)
      for (int j = i + 1; j < n; j++) {
        bool same = true;
        for (int d = 0; d < dim && same; d++) same = std::fabs(pts[(size_t)i * dim + d] - pts[(size_t)j * dim + d]) < 1e-12;
        if (same) dup[j] = 1;
      }
    std::vector<double> out;
    for (int i = 0; i < n; i++) if (!dup[i]) out.insert(out.end(), pts.begin() + (size_t)i * dim, pts.begin() + (size_t)(i + 1) * dim);
    pts.swap(out);
  }
  return g_nodes[key] = pts;
}

}  // namespace

const std::vector<double>& node_set(int dim, int order, Geometry g) {
  std::lock_guard<std::mutex> lk(g_mutex);
  return node_set_locked(dim, order, g);
}

CubatureRule make_cubature(int dim, int degree, Geometry g) {
  static const int maxDeg[4] = {20, 20, 20, 10};
  if (dim > 3 || dim < 0) throw std::runtime_error("Cubature : Constructor : the requested space dimension (" + std::to_string(dim) + ") is too large and thus not implemented in the cubature rules yet. (maxDim = 3)");
  if (degree > maxDeg[dim]) throw std::runtime_error("Cubature : Constructor : the requested polynomial order (" + std::to_string(degree) + ") is too large for dimension " + std::to_string(dim) + " and thus not implemented in the cubature rules yet. (maxOrder = " + std::to_string(maxDeg[dim]) + ")");
  CubatureRule r;
  r.dim = dim; r.degree = degree;
  if (dim == 0) return r;
  int nip = -1;
  for (const auto& e : kHfxNipMap) if (e.dim == dim && e.degree == degree && e.geom == (int)g) nip = e.nip;
  if (nip < 0) throw std::runtime_error("Cubature : determineRule : no rule for this (dimension, degree, geometry)");
  for (const auto& e : kHfxRules)
    if (e.dim == dim && e.nip == nip && e.geom == (int)g) {
      r.nIP = nip;
      r.coords.resize((size_t)nip * dim); r.weights.resize(nip);
      const double* p = kHfxRuleData + e.offset;
      for (int i = 0; i < nip; i++) { for (int d = 0; d < dim; d++) r.coords[(size_t)i * dim + d] = p[(size_t)i * (dim + 1) + d]; r.weights[i] = p[(size_t)i * (dim + 1) + dim]; }
      return r;
    }
  throw std::runtime_error("Cubature : determineRule : rule table missing");
}

static Geometry parse_geom(const std::string& s) {
  if (s == "simplex") return kSimplex;
  if (s == "orthotope" || s == "quad" || s == "hex") return kOrthotope;
  throw std::runtime_error("ReferenceElement : setGeometry : Element type " + s + " is not yet supported.");
}

RefElement::RefElement(int dim, int order, const std::string& geom) : RefElement(dim, order, parse_geom(geom)) {}

RefElement::RefElement(int dim, int order, Geometry g) : dim_(dim), order_(order), geom_(g) {
  if (dim > 3) throw std::runtime_error("ReferenceElement : setDim : The spatial dimension " + std::to_string(dim) + " is not yet supported.");
  if (dim < 0) throw std::runtime_error("ReferenceElement : setDim : The spatial dimension " + std::to_string(dim) + " is negative.");
  if (order > max_order(dim, g)) throw std::runtime_error("ReferenceElement : setOrder : The interpolation order " + std::to_string(order) + " is not yet supported for dimension " + std::to_string(dim) + ".");
  if (order < 0) throw std::runtime_error("ReferenceElement : setOrder : The interpolation order " + std::to_string(order) + " is negative.");
  build();
}

RefElement::~RefElement() { delete face_; }

void RefElement::build() {
  nodes_ = node_set(dim_, order_, geom_);
  nN_ = dim_ == 0 ? 1 : (int)nodes_.size() / dim_;
  cub_ = make_cubature(dim_, geom_ == kSimplex ? 2 * order_ : 4 * order_, geom_);
  nFc_ = geom_ == kSimplex ? dim_ + 1 : 2 * dim_;
  if (dim_ > 0) face_ = new RefElement(dim_ - 1, order_, geom_);
  // face-node maps
  if (dim_ == 1 && order_ > 0) {
    faceNodes_ = {0, 1};
    for (int i = 2; i <= order_; i++) innerNodes_.push_back(i);
  } else if (dim_ > 1 && order_ > 0) {
    const Topo& tp = topo(dim_, geom_);
    int q = dim_ - 1, nNf = face_->numNodes();
    std::vector<double> x(dim_);
    int mx = -1;
    for (auto& ent : tp.ents) {
      if ((int)ent.size() != (int)topo(q, geom_).v.size()) continue;
      Embed em(dim_, q, geom_, ent);
      for (int l = 0; l < nNf; l++) {
        em.map(&face_->nodes()[(size_t)l * q], em.xi0.data(), x.data());
        int found = -1;
        for (int k = 0; k < nN_ && found < 0; k++) {
          bool eq = true;
          for (int d = 0; d < dim_ && eq; d++) eq = !(std::fabs(nodes_[(size_t)k * dim_ + d] - x[d]) > 1e-12);
          if (eq) found = k;
        }
        if (found < 0) throw std::runtime_error("ReferenceElement : determineFaceNodes : could not find one of the face nodes.");
        faceNodes_.push_back(found);
        mx = std::max(mx, found);
      }
    }
    for (int i = mx + 1; i < nN_; i++) innerNodes_.push_back(i);
  }
  if (dim_ == 0) return;
  // node -> mode multi-indices
  auto combos = index_tuples(0, order_, dim_);
  for (auto& c : combos) {
    int s = 0; for (int v : c) s += v;
    if (geom_ == kSimplex && s > order_) continue;
    modeMap_.insert(modeMap_.end(), c.begin(), c.end());
  }
  if ((int)modeMap_.size() != nN_ * dim_) throw std::runtime_error("ReferenceElement : determineNodeToModeMap : mode/node count mismatch");
  std::vector<double> V((size_t)nN_ * nN_);
  for (int i = 0; i < nN_; i++) { auto m = modes(&nodes_[(size_t)i * dim_]); for (int j = 0; j < nN_; j++) V[(size_t)i * nN_ + j] = m[j]; }
  invV_ = invert(V, nN_);
  int nIP = cub_.nIP;
  ipShape_.resize((size_t)nIP * nN_); ipDShape_.resize((size_t)nIP * nN_ * dim_);
  for (int ip = 0; ip < nIP; ip++) {
    auto s = interpolate(&cub_.coords[(size_t)ip * dim_]);
    auto d = interpolateDeriv(&cub_.coords[(size_t)ip * dim_]);
    for (int i = 0; i < nN_; i++) ipShape_[(size_t)ip * nN_ + i] = s[i];
    for (int i = 0; i < nN_ * dim_; i++) ipDShape_[(size_t)ip * nN_ * dim_ + i] = d[i];
  }
}

// collapsed coordinates of the simplex
static void collapse(int dim, const double* c, double* m) {
  if (dim == 1) { m[0] = c[0]; return; }
  if (dim == 2) { m[0] = (c[1] != 1.0) ? 2.0 * (1.0 + c[0]) / (1.0 - c[1]) - 1.0 : -1.0; m[1] = c[1]; return; }
  m[0] = ((c[1] + c[2]) != 0.0) ? -2.0 * (1.0 + c[0]) / (c[1] + c[2]) - 1.0 : -1.0;
  m[1] = (c[2] != 1.0) ? 2.0 * (1.0 + c[1]) / (1.0 - c[2]) - 1.0 : -1.0;
  m[2] = c[2];
}

// value (deriv=false) or derivative w.r.t. the collapsed coordinate (deriv=true) of the l-th 1-D factor of mode `md`
static double factor(Geometry g, int l, const int* md, const double* mc, bool deriv) {
  if (g == kOrthotope || l == 0) return deriv ? jacobiDP(md[l], 0, 0, mc[l]) : jacobiP(md[l], 0, 0, mc[l]);
  if (l == 1) {
    double a = 2.0 * md[0] + 1.0, s2 = std::sqrt(2.0);
    if (!deriv) return s2 * jacobiP(md[1], a, 0, mc[1]) * ipow(1 - mc[1], md[0]);
    int pw = md[0] - 1 < 0 ? 0 : md[0] - 1;
    return s2 * jacobiDP(md[1], a, 0, mc[1]) * ipow(1 - mc[1], md[0]) + s2 * jacobiP(md[1], a, 0, mc[1]) * (-md[0] * ipow(1 - mc[1], pw));
  }
  int e = md[1] + md[0];
  double a = 2.0 * (e + 1.0);
  if (!deriv) return 2.0 * jacobiP(md[2], a, 0, mc[2]) * ipow(1 - mc[2], e);
  int pw = e - 1 < 0 ? 0 : e - 1;
  return 2.0 * jacobiDP(md[2], a, 0, mc[2]) * ipow(1 - mc[2], e) + 2.0 * jacobiP(md[2], a, 0, mc[2]) * (-e * ipow(1 - mc[2], pw));
}

std::vector<double> RefElement::modes(const double* pt) const {
  std::vector<double> r(nN_, 1.0);
  double mc[3];
  if (geom_ == kSimplex) collapse(dim_, pt, mc); else for (int d = 0; d < dim_; d++) mc[d] = pt[d];
  for (int j = 0; j < nN_; j++) for (int k = 0; k < dim_; k++) r[j] *= factor(geom_, k, &modeMap_[(size_t)j * dim_], mc, false);
  return r;
}

std::vector<double> RefElement::derivModes(const double* pt) const {
  std::vector<double> r((size_t)nN_ * dim_, 1.0);
  double mc[3];
  double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // J[i][j] = d(collapsed_j)/d(pt_i)
  if (geom_ == kOrthotope) { for (int d = 0; d < dim_; d++) { mc[d] = pt[d]; J[d * dim_ + d] = 1.0; } }
  else {
    collapse(dim_, pt, mc);
    const double eps = 1e-6;  // guard at the collapsed vertex, as the reference
    if (dim_ == 1) J[0] = 1.0;
    else if (dim_ == 2) {
      if (pt[1] != 1.0) { J[0] = 2.0 / (1.0 - pt[1]); J[2] = 2.0 * (1.0 + pt[0]) * (1.0 / std::pow(1.0 - pt[1], 2.0)); }
      else { J[0] = 2.0 / eps; J[2] = 0.0; }
      J[3] = 1.0;
    } else {
      if ((pt[1] + pt[2]) != 0.0) {
        J[0] = -2.0 / (pt[1] + pt[2]);
        J[3] = 2.0 * (1.0 + pt[0]) * (1.0 / std::pow(pt[1] + pt[2], 2.0));
        J[6] = J[3];
      } else J[0] = -2.0 / eps;
      if (pt[2] != 1.0) { J[4] = 2.0 / (1.0 - pt[2]); J[7] = 2.0 * (1.0 + pt[1]) * (1.0 / std::pow(1.0 - pt[2], 2.0)); }
      else J[4] = 2.0 / eps;
      J[8] = 1.0;
    }
  }
  for (int j = 0; j < nN_; j++) {
    const int* md = &modeMap_[(size_t)j * dim_];
    double g[3] = {1, 1, 1};
    for (int k = 0; k < dim_; k++) for (int l = 0; l < dim_; l++) g[k] *= factor(geom_, l, md, mc, l == k);
    for (int i = 0; i < dim_; i++) { double s = 0; for (int k = 0; k < dim_; k++) s += J[i * dim_ + k] * g[k]; r[(size_t)j * dim_ + i] = s; }
  }
  return r;
}

std::vector<double> RefElement::interpolate(const double* pt) const {
  if (dim_ == 0) return std::vector<double>(nN_, 1.0);
  auto m = modes(pt);
  std::vector<double> r(nN_, 0.0);
  for (int k = 0; k < nN_; k++) { double mk = m[k]; for (int i = 0; i < nN_; i++) r[i] += mk * invV_[(size_t)k * nN_ + i]; }
  return r;
}

std::vector<double> RefElement::interpolateDeriv(const double* pt) const {
  if (dim_ == 0) return std::vector<double>(nN_, 0.0);
  auto dm = derivModes(pt);
  std::vector<double> r((size_t)nN_ * dim_, 0.0);
  for (int d = 0; d < dim_; d++)
    for (int k = 0; k < nN_; k++) { double mk = dm[(size_t)k * dim_ + d]; for (int i = 0; i < nN_; i++) r[(size_t)i * dim_ + d] += mk * invV_[(size_t)k * nN_ + i]; }
  return r;
}

}  // namespace hfx
