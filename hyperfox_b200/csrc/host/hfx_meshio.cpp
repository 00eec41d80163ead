// Gmsh reader + order-p mesh generation without MOAB.  Reference: tools/convertGmsh2H5HO.cpp:117-257 (generateHigherOrderMesh),
// :259-363 (readMesh), :366-397 (generateCellNodes).  Numbering convention of the intermediate entities (what MOAB decides there):
//   * entities present in the input file come first, in file order and with the file's vertex order;
//   * the missing sub-entities are created while walking the cells in ascending id and, inside a cell, the sub-entities in the
//     canonical order below; a new one takes the next id and the vertex order of that first appearance;
//   * the sub-entities adjacent to a cell are visited in ascending id.
#include "hfx_meshio.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iterator>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>

#include "hfx_refel.h"

namespace hfx {

namespace {
const int kEdges2[3][2] = {{0, 1}, {1, 2}, {2, 0}};
const int kEdges3[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};
const int kTetFaces[4][3] = {{0, 1, 3}, {1, 2, 3}, {0, 3, 2}, {0, 2, 1}};

[[noreturn]] void fail(const char* fn, const std::string& msg) { throw std::runtime_error(std::string("MeshIo : ") + fn + " : " + msg); }

// inverse of a small (n <= 3) row-major matrix by Gauss-Jordan with partial pivoting
void small_inverse(int n, const double* a, double* inv) {
  double m[3][6];
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { m[i][j] = a[i * n + j]; m[i][n + j] = i == j ? 1.0 : 0.0; }
  for (int k = 0; k < n; k++) {
    int pv = k;
    for (int i = k + 1; i < n; i++) if (std::fabs(m[i][k]) > std::fabs(m[pv][k])) pv = i;
    if (m[pv][k] == 0.0) fail("generateCellNodes", "degenerate cell");
    if (pv != k) for (int j = 0; j < 2 * n; j++) std::swap(m[k][j], m[pv][j]);
    const double d = 1.0 / m[k][k];
    for (int j = 0; j < 2 * n; j++) m[k][j] *= d;
    for (int i = 0; i < n; i++) if (i != k) { const double f = m[i][k]; for (int j = 0; j < 2 * n; j++) m[i][j] -= f * m[k][j]; }
  }
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) inv[i * n + j] = m[i][n + j];
}

// generateCellNodes: node = T (xi - xi_0) + x_0 with T = [x_i - x_0] [xi_i - xi_0]^-1
void cell_nodes(const RefElement& re, int dim, const double* lin /*[td+1][dim]*/, std::vector<double>* out) {
  const int td = re.dim(), nN = re.numNodes();
  const std::vector<double>& ref = re.nodes();
  double locT[9], locTi[9], X[9], T[9];
  for (int i = 0; i < td; i++) for (int j = 0; j < td; j++) locT[j * td + i] = ref[(size_t)(i + 1) * td + j] - ref[j];
  small_inverse(td, locT, locTi);
  for (int i = 0; i < td; i++) for (int j = 0; j < dim; j++) X[j * td + i] = lin[(i + 1) * dim + j] - lin[j];
  for (int j = 0; j < dim; j++) for (int r = 0; r < td; r++) { double s = 0.0; for (int k = 0; k < td; k++) s += X[j * td + k] * locTi[k * td + r]; T[j * td + r] = s; }
  out->resize((size_t)nN * dim);
  for (int i = 0; i < nN; i++)
    for (int j = 0; j < dim; j++) {
      double s = 0.0;
      for (int r = 0; r < td; r++) s += T[j * td + r] * (ref[(size_t)i * td + r] - ref[r]);
      (*out)[(size_t)i * dim + j] = s + lin[j];
    }
}
}  // namespace

void read_msh(const std::string& path, MshFile* out) {
  std::ifstream f(path);
  if (!f) fail("readMsh", "could not load mesh file: " + path);
  std::string line;
  std::vector<long long> tags;
  std::vector<double> raw;
  std::vector<long long> el[4];
  auto trimmed = [](std::string s) { while (!s.empty() && (s.back() == '\r' || s.back() == ' ')) s.pop_back(); return s; };
  while (std::getline(f, line)) {
    line = trimmed(line);
    if (line == "$MeshFormat") {
      std::getline(f, line);
      std::istringstream is(line);
      std::string ver; int type = -1;
      is >> ver >> type;
      if (ver.empty() || ver[0] != '2' || type != 0) fail("readMsh", "only the Gmsh 2.x ASCII format is supported");
    } else if (line == "$Nodes") {
      long long n = 0;
      f >> n;
      tags.resize((size_t)n); raw.resize((size_t)n * 3);
      for (long long k = 0; k < n; k++) f >> tags[(size_t)k] >> raw[(size_t)k * 3] >> raw[(size_t)k * 3 + 1] >> raw[(size_t)k * 3 + 2];
      if (!f) fail("readMsh", "truncated $Nodes section");
    } else if (line == "$Elements") {
      long long n = 0;
      f >> n;
      for (long long k = 0; k < n; k++) {
        long long id; int ty, ntags;
        f >> id >> ty >> ntags;
        for (int t = 0; t < ntags; t++) { long long tag; f >> tag; }
        int td = -1;
        if (ty == 15) td = 0; else if (ty == 1) td = 1; else if (ty == 2) td = 2; else if (ty == 4) td = 3;
        if (td < 0) fail("readMsh", "element type " + std::to_string(ty) + " is not supported (linear simplices only)");
        for (int v = 0; v <= td; v++) { long long nd; f >> nd; if (td > 0) el[td].push_back(nd); }
      }
      if (!f) fail("readMsh", "truncated $Elements section");
    }
  }
  if (tags.empty()) fail("readMsh", "no $Nodes section in " + path);
  std::vector<size_t> order(tags.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return tags[a] < tags[b]; });
  std::map<long long, int> remap;
  out->nodes.resize(raw.size());
  for (size_t i = 0; i < order.size(); i++) {
    remap[tags[order[i]]] = (int)i;
    for (int d = 0; d < 3; d++) out->nodes[i * 3 + d] = raw[order[i] * 3 + d];
  }
  for (int k = 1; k <= 3; k++) {
    out->elems[k].resize(el[k].size());
    for (size_t i = 0; i < el[k].size(); i++) {
      auto it = remap.find(el[k][i]);
      if (it == remap.end()) fail("readMsh", "an element refers to an unknown node tag");
      out->elems[k][i] = it->second;
    }
  }
}

void high_order_mesh(int dim, int order, int nLin, const double* lin, int nCells, const int* cells, const std::vector<int> (&existing)[4],
                     std::vector<double>* nodes, std::vector<int>* hoCells) {
  if (dim != 2 && dim != 3) fail("generateHigherOrderMesh", "dimension must be 2 or 3");
  // sub-entities of every topological dimension k < dim: connectivity + ascending ids per cell
  std::vector<int> conn[4];
  std::vector<int> adj[4];
  int nSub[4] = {0, 0, 0, 0};
  for (int k = 1; k < dim; k++) {
    const int nv = k + 1;
    const int ns = k == 1 ? (dim == 2 ? 3 : 6) : 4;
    nSub[k] = ns;
    std::map<std::array<int, 3>, int> ids;
    auto key = [&](const int* v) { std::array<int, 3> a = {-1, -1, -1}; for (int i = 0; i < nv; i++) a[i] = v[i]; std::sort(a.begin(), a.begin() + nv); return a; };
    conn[k] = existing[k];
    for (size_t e = 0; e * nv < existing[k].size(); e++) ids.emplace(key(&existing[k][e * nv]), (int)e);
    adj[k].resize((size_t)nCells * ns);
    for (int c = 0; c < nCells; c++) {
      for (int s = 0; s < ns; s++) {
        int v[3];
        for (int i = 0; i < nv; i++) {
          const int loc = k == 1 ? (dim == 2 ? kEdges2[s][i] : kEdges3[s][i]) : kTetFaces[s][i];
          v[i] = cells[(size_t)c * (dim + 1) + loc];
        }
        auto ins = ids.emplace(key(v), (int)(conn[k].size() / nv));
        if (ins.second) conn[k].insert(conn[k].end(), v, v + nv);
        adj[k][(size_t)c * ns + s] = ins.first->second;
      }
      std::sort(adj[k].begin() + (size_t)c * ns, adj[k].begin() + (size_t)(c + 1) * ns);
    }
  }
  conn[dim].assign(cells, cells + (size_t)nCells * (dim + 1));
  nSub[dim] = 1;
  adj[dim].resize(nCells);
  for (int c = 0; c < nCells; c++) adj[dim][c] = c;

  std::vector<std::unique_ptr<RefElement>> re(dim + 1);
  for (int k = 1; k <= dim; k++) re[k].reset(new RefElement(k, order, kSimplex));
  const int nN = re[dim]->numNodes();
  nodes->clear();
  hoCells->assign((size_t)nCells * nN, -1);
  std::vector<int> vertId(nLin, -1);
  std::vector<std::vector<int>> entFirst(dim + 1);     // first generated node of an entity, -1 if not generated yet
  for (int k = 1; k <= dim; k++) entFirst[k].assign(conn[k].size() / (k + 1), -1);
  std::vector<double> elNodes, subNodes, linEl((size_t)(dim + 1) * dim), linSub((size_t)(dim + 1) * dim);
  int next = 0;
  for (int e = 0; e < nCells; e++) {
    for (int i = 0; i <= dim; i++) {
      const int v = cells[(size_t)e * (dim + 1) + i];
      if (v < 0 || v >= nLin) fail("generateHigherOrderMesh", "cell refers to a vertex outside the node list");
      for (int d = 0; d < dim; d++) linEl[(size_t)i * dim + d] = lin[(size_t)v * dim + d];
      if (vertId[v] < 0) { vertId[v] = next++; nodes->insert(nodes->end(), &lin[(size_t)v * dim], &lin[(size_t)v * dim] + dim); }
      (*hoCells)[(size_t)e * nN + i] = vertId[v];
    }
    cell_nodes(*re[dim], dim, linEl.data(), &elNodes);
    for (int k = 1; k <= dim; k++) {
      const std::vector<int>& inner = re[k]->innerNodes();
      const int nIn = (int)inner.size();
      if (nIn == 0) continue;
      for (int s = 0; s < nSub[k]; s++) {
        const int cid = adj[k][(size_t)e * nSub[k] + s];
        if (entFirst[k][cid] < 0) {
          for (int i = 0; i <= k; i++) for (int d = 0; d < dim; d++) linSub[(size_t)i * dim + d] = lin[(size_t)conn[k][(size_t)cid * (k + 1) + i] * dim + d];
          cell_nodes(*re[k], dim, linSub.data(), &subNodes);
          entFirst[k][cid] = next;
          for (int l = 0; l < nIn; l++) { nodes->insert(nodes->end(), &subNodes[(size_t)inner[l] * dim], &subNodes[(size_t)inner[l] * dim] + dim); next++; }
        }
        for (int l = 0; l < nIn; l++) {
          const int nid = entFirst[k][cid] + l;
          int hit = -1;
          for (int n = 0; n < nN && hit < 0; n++) {
            bool eq = true;
            for (int d = 0; d < dim && eq; d++) eq = std::fabs((*nodes)[(size_t)nid * dim + d] - elNodes[(size_t)n * dim + d]) < 1e-8;
            if (eq) hit = n;
          }
          if (hit < 0) fail("generateHigherOrderMesh", "one of the cell nodes could not be found in element");
          (*hoCells)[(size_t)e * nN + hit] = nid;
        }
      }
    }
  }
}

// ---- HDF5 subset reader -------------------------------------------------------------------------------------------------------
namespace {
struct H5File {
  std::vector<unsigned char> b;
  [[noreturn]] static void bad(const std::string& m) { throw std::runtime_error("HDF5Io : loadMesh : " + m); }
  // every offset / size below comes from the (untrusted) file: range checks are written so that they cannot wrap around
  bool inside(uint64_t off, uint64_t n) const { return off <= b.size() && n <= b.size() - off; }
  mutable size_t visited = 0;
  void visit() const { if (++visited > 1000000) bad("corrupt file (cyclic or oversized header structure)"); }
  uint64_t u(size_t off, int n) const {
    if (!inside(off, (uint64_t)n)) bad("truncated file");
    uint64_t v = 0;
    for (int i = n - 1; i >= 0; i--) v = (v << 8) | b[off + i];
    return v;
  }
  bool tag(size_t off, const char* t) const { return inside(off, 4) && std::equal(t, t + 4, b.begin() + off); }
  struct Entry { uint64_t nameOff = 0, ohdr = 0, btree = 0, heap = 0; bool cached = false; };
  struct Msg { int type; size_t data, size; };
  Entry entry(size_t off) const {
    Entry e;
    e.nameOff = u(off, 8); e.ohdr = u(off + 8, 8);
    if (u(off + 16, 4) == 1) { e.cached = true; e.btree = u(off + 24, 8); e.heap = u(off + 32, 8); }
    return e;
  }
  std::vector<Msg> messages(size_t ohdr) const {
    if (u(ohdr, 1) != 1) bad("only version-1 object headers are supported");
    const size_t nmsg = u(ohdr + 2, 2);
    std::vector<std::pair<size_t, size_t>> blocks = {{ohdr + 16, (size_t)u(ohdr + 8, 4)}};
    std::vector<Msg> out;
    for (size_t bi = 0; bi < blocks.size() && out.size() < nmsg; bi++) {
      visit();
      size_t off = blocks[bi].first;
      if (!inside(off, blocks[bi].second)) bad("object header block outside the file");
      const size_t end = off + blocks[bi].second;
      while (off + 8 <= end && out.size() < nmsg) {
        visit();
        const int type = (int)u(off, 2);
        const size_t sz = u(off + 2, 2), data = off + 8;
        if (!inside(data, sz)) bad("object header message outside the file");
        if (type == 0x10) blocks.push_back({(size_t)u(data, 8), (size_t)u(data + 8, 8)});   // continuation block
        out.push_back({type, data, sz});
        off = data + sz;
      }
    }
    return out;
  }
  void walk(size_t node, size_t heapData, std::map<std::string, Entry>* out, int depth) const {
    visit();
    if (!tag(node, "TREE") || depth > 16) bad("corrupt group B-tree");
    const int level = (int)u(node + 5, 1), nent = (int)u(node + 6, 2);
    const size_t p = node + 8 + 16;
    for (int i = 0; i < nent; i++) {
      const size_t child = u(p + 8 + (size_t)i * 16, 8);
      if (level > 0) { walk(child, heapData, out, depth + 1); continue; }
      if (!tag(child, "SNOD")) bad("corrupt symbol table node");
      const int ns = (int)u(child + 6, 2);
      for (int k = 0; k < ns; k++) {
        visit();
        const Entry e = entry(child + 8 + (size_t)k * 40);
        if (!inside(heapData, e.nameOff)) bad("symbol name outside the file");
        size_t s = heapData + e.nameOff;
        std::string name;
        while (s < b.size() && b[s]) name.push_back((char)b[s++]);
        (*out)[name] = e;
      }
    }
  }
  std::map<std::string, Entry> children(Entry g) const {
    if (!g.cached) for (const Msg& m : messages(g.ohdr)) if (m.type == 0x11) { g.btree = u(m.data, 8); g.heap = u(m.data + 8, 8); g.cached = true; }
    if (!g.cached || !tag(g.heap, "HEAP")) bad("not a symbol-table group");
    std::map<std::string, Entry> out;
    walk(g.btree, u(g.heap + 8 + 16, 8), &out, 0);
    return out;
  }
  // dataset -> shape, element class (0 integer, 1 float), element size, address of the contiguous data
  void dataset(const Entry& d, std::vector<uint64_t>* shape, int* cls, int* size, size_t* addr) const {
    bool haveS = false, haveT = false, haveL = false;
    for (const Msg& m : messages(d.ohdr)) {
      if (m.type == 0x1) {
        const int v = (int)u(m.data, 1), rank = (int)u(m.data + 1, 1);
        const size_t base = v == 1 ? m.data + 8 : m.data + 4;
        shape->clear();
        for (int i = 0; i < rank; i++) shape->push_back(u(base + 8 * (size_t)i, 8));
        haveS = true;
      } else if (m.type == 0x3) {
        *cls = (int)(u(m.data, 1) & 0x0F); *size = (int)u(m.data + 4, 4); haveT = true;
      } else if (m.type == 0x8) {
        if (u(m.data, 1) != 3 || u(m.data + 1, 1) != 1) bad("only contiguous dataset layouts are supported");
        *addr = (size_t)u(m.data + 2, 8); haveL = true;
        // (HADDR_UNDEF: no storage allocated -- legal for an empty dataset; the callers check the address against the bytes they need)
      }
    }
    if (!haveS || !haveT || !haveL) bad("incomplete dataset header");
  }
};
}  // namespace

void read_h5_mesh(const std::string& path, H5Mesh* out) {
  H5File f;
  {
    std::ifstream in(path, std::ios::binary);
    if (!in) H5File::bad("could not open " + path);
    f.b.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
  }
  static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
  if (f.b.size() < 96 || !std::equal(sig, sig + 8, f.b.begin())) H5File::bad(path + " is not an HDF5 file");
  if (f.b[8] != 0 || f.b[13] != 8 || f.b[14] != 8) H5File::bad("only superblock version 0 with 8-byte offsets is supported");
  const H5File::Entry root = f.entry(24 + 4 * 8);
  auto top = f.children(root);
  if (!top.count("Mesh")) H5File::bad("the file has no Mesh group");
  auto mesh = f.children(top["Mesh"]);
  if (!mesh.count("Nodes") || !mesh.count("Cells")) H5File::bad("the Mesh group must hold the Nodes and Cells datasets");
  std::vector<uint64_t> shape;
  int cls = -1, size = 0;
  size_t addr = 0;
  f.dataset(mesh["Nodes"], &shape, &cls, &size, &addr);
  if (shape.size() != 2 || cls != 1 || size != 8) H5File::bad("Nodes must be a two-dimensional float64 dataset");
  if (shape[1] == 0 || shape[1] > 3 || shape[0] > f.b.size() / (8 * shape[1])) H5File::bad("truncated Nodes dataset");
  const size_t nn = (size_t)(shape[0] * shape[1]);
  if (!f.inside(addr, (uint64_t)nn * 8)) H5File::bad(addr == UINT64_MAX ? "dataset without allocated storage (HADDR_UNDEF)" : "truncated Nodes dataset");
  out->dimNodeSpace = (int)shape[1];
  out->nodes.resize(nn);
  std::memcpy(out->nodes.data(), f.b.data() + addr, nn * 8);
  f.dataset(mesh["Cells"], &shape, &cls, &size, &addr);
  if (shape.size() != 2 || cls != 0 || (size != 4 && size != 8)) H5File::bad("Cells must be a two-dimensional integer dataset");
  if (shape[1] == 0 || shape[1] > 4096 || shape[0] > f.b.size() / ((uint64_t)size * shape[1])) H5File::bad("truncated Cells dataset");
  const size_t nc = (size_t)(shape[0] * shape[1]);
  if (!f.inside(addr, (uint64_t)nc * size)) H5File::bad(addr == UINT64_MAX ? "dataset without allocated storage (HADDR_UNDEF)" : "truncated Cells dataset");
  out->nodesPerCell = (int)shape[1];
  out->cells.resize(nc);
  for (size_t i = 0; i < nc; i++) out->cells[i] = (int)(int64_t)f.u(addr + i * size, size);
}


namespace {
H5File open_h5(const std::string& path, const char* who) {
  H5File f;
  std::ifstream in(path, std::ios::binary);
  if (!in) throw std::runtime_error(std::string("HDF5Io : ") + who + " : could not open " + path);
  f.b.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
  static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
  if (f.b.size() < 96 || !std::equal(sig, sig + 8, f.b.begin())) throw std::runtime_error(std::string("HDF5Io : ") + who + " : the file " + path + " is not an hdf5 file.");
  if (f.b[8] != 0 || f.b[13] != 8 || f.b[14] != 8) H5File::bad("only superblock version 0 with 8-byte offsets is supported");
  return f;
}
}  // namespace

bool h5_has_mesh(const std::string& path) {
  H5File f = open_h5(path, "load");
  return f.children(f.entry(24 + 4 * 8)).count("Mesh") > 0;
}

std::vector<std::string> h5_field_names(const std::string& path) {
  H5File f = open_h5(path, "loadFields");
  auto top = f.children(f.entry(24 + 4 * 8));
  std::vector<std::string> out;
  if (!top.count("FieldData")) return out;
  for (const auto& kv : f.children(top["FieldData"])) out.push_back(kv.first);
  return out;
}

void read_h5_field(const std::string& path, const std::string& name, H5Field* out) {
  auto fail = [&](const std::string& m) { throw std::runtime_error("HDF5Io : loadFields : " + m); };
  H5File f = open_h5(path, "loadFields");
  auto top = f.children(f.entry(24 + 4 * 8));
  if (!top.count("FieldData")) fail("the file has no FieldData group");
  auto fd = f.children(top["FieldData"]);
  if (!fd.count(name)) fail("field with name " + name + " was not found in FieldData");
  std::vector<uint64_t> shape;
  int cls = -1, size = 0;
  size_t addr = 0;
  f.dataset(fd[name], &shape, &cls, &size, &addr);
  if (shape.size() != 3 || cls != 1 || size != 8) fail("a field must be a three-dimensional float64 dataset");
  for (int k = 0; k < 3; k++) if (shape[k] > f.b.size()) fail("could not get shape of field dataspace");
  const uint64_t n01 = shape[0] * shape[1];
  if ((shape[1] && n01 / shape[1] != shape[0]) || n01 > f.b.size() || (shape[2] && n01 * shape[2] > f.b.size() / 8)) fail("truncated field dataset");
  const size_t n = (size_t)(n01 * shape[2]);
  if (n && !f.inside(addr, (uint64_t)n * 8)) fail("could not load field values into field");
  out->name = name;
  for (int k = 0; k < 3; k++) out->shape[k] = (long long)shape[k];
  out->vals.resize(n);
  if (n) std::memcpy(out->vals.data(), f.b.data() + addr, n * 8);
  // attribute "ftype" (version-1 attribute message: name, datatype, dataspace, each padded to 8 bytes, then the value; any integer width)
  bool found = false;
  for (const H5File::Msg& m : f.messages(fd[name].ohdr)) {
    if (m.type != 0xc || m.size < 8) continue;
    if (f.u(m.data, 1) != 1) fail("only version-1 attribute messages are supported");
    const size_t ns = f.u(m.data + 2, 2), ts = f.u(m.data + 4, 2), ss = f.u(m.data + 6, 2);
    auto pad8 = [](size_t x) { return (x + 7) & ~(size_t)7; };
    const size_t pn = m.data + 8, pt = pn + pad8(ns), ps = pt + pad8(ts), pv = ps + pad8(ss);
    if (!f.inside(pn, ns) || ns < 1 || pv > m.data + m.size) fail("corrupt attribute message");
    std::string an(reinterpret_cast<const char*>(f.b.data() + pn), ns - 1);
    if (an != "ftype") continue;
    if ((f.u(pt, 1) & 0x0F) != 0) fail("attribute ftype must be an integer");
    const int isz = (int)f.u(pt + 4, 4);
    if ((isz != 1 && isz != 2 && isz != 4 && isz != 8) || pv + isz > m.data + m.size) fail("corrupt attribute message");
    const uint64_t raw = f.u(pv, isz);
    out->ftype = (int)(int64_t)(isz == 8 ? raw : (isz == 4 ? (uint64_t)(int64_t)(int32_t)raw : (isz == 2 ? (uint64_t)(int64_t)(int16_t)raw : (uint64_t)(int64_t)(int8_t)raw)));
    found = true;
  }
  if (!found) fail("attribute ftype could not be found for field: " + name);
}

// ---- HDF5 subset writer ------------------------------------------------------------------------------------------------------------------------------------
namespace {
struct H5Writer {
  std::vector<unsigned char> b;
  uint64_t eoa = 96;                                   // the superblock
  struct Aggr { uint64_t addr = 0, end = 0; bool open = false; } meta, sdata;
  static constexpr uint64_t kBlock = 2048;
  void put(uint64_t off, uint64_t v, int n) { if (b.size() < off + n) b.resize(off + n, 0); for (int i = 0; i < n; i++) b[off + i] = (unsigned char)(v >> (8 * i)); }
  void bytes(uint64_t off, const void* p, size_t n) { if (b.size() < off + n) b.resize(off + n, 0); if (n) std::memcpy(b.data() + off, p, n); }
  // libhdf5's block aggregators (H5MFaggr.c) as far as these files exercise them: bump inside the block; a block that ends at the end of the file grows by what is
  // missing; otherwise a new block of max(2 KB, size) starts at the end of the file
  uint64_t aggr(Aggr& a, uint64_t size) {
    if (a.open && a.addr + size <= a.end) { const uint64_t r = a.addr; a.addr += size; return r; }
    if (a.open && a.end == eoa) { eoa += size - (a.end - a.addr); a.end = eoa; const uint64_t r = a.addr; a.addr += size; return r; }
    a.open = true; a.addr = eoa; a.end = eoa + std::max(kBlock, size); eoa = a.end;
    const uint64_t r = a.addr; a.addr += size; return r;
  }
  uint64_t allocMeta(uint64_t size) { return aggr(meta, size); }
  uint64_t allocRaw(uint64_t size) { if (size < kBlock) return aggr(sdata, size); const uint64_t r = eoa; eoa += size; return r; }
  void close() {   // blocks that end at the end of the file give their unused tail back
    for (int pass = 0; pass < 2; pass++) {
      if (sdata.open && sdata.end == eoa) { eoa = sdata.addr; sdata.end = sdata.addr; }
      if (meta.open && meta.end == eoa) { eoa = meta.addr; meta.end = meta.addr; }
    }
  }
  struct Group {
    uint64_t ohdr = 0, btree = 0, heap = 0, heapData = 0, heapSize = 0, heapUsed = 8;
    std::vector<uint64_t> snods;
    struct Child { std::string name; uint64_t nameOff, ohdr; bool isGroup; uint64_t btree, heap; };
    std::vector<Child> children;
  };
  // H5Gcreate: object header (one symbol-table message), B-tree node of the group (K = 16: 544 bytes), local heap (32-byte header + data segment)
  Group makeGroup(size_t nameBytes) {
    Group g;
    g.ohdr = allocMeta(40);
    g.btree = allocMeta(544);
    g.heapSize = std::max<uint64_t>(88, ((8 + nameBytes + 16 + 7) / 8) * 8);     // libhdf5 starts with 88 bytes and would grow the segment; sized at once here
    g.heap = allocMeta(32 + g.heapSize);
    g.heapData = g.heap + 32;
    return g;
  }
  void link(Group& g, const std::string& name, uint64_t ohdr, bool isGroup, uint64_t btree, uint64_t heap) {
    if (g.children.size() % 8 == 0) g.snods.push_back(allocMeta(328));        // symbol-table nodes hold 2K = 8 entries
    const uint64_t len = ((name.size() + 1 + 7) / 8) * 8;
    g.children.push_back({name, g.heapUsed, ohdr, isGroup, btree, heap});
    g.heapUsed += len;
  }
  void finishGroup(const Group& g) {
    put(g.ohdr, 1, 1); put(g.ohdr + 2, 1, 2); put(g.ohdr + 4, 1, 4); put(g.ohdr + 8, 24, 4);
    put(g.ohdr + 16, 0x11, 2); put(g.ohdr + 18, 16, 2); put(g.ohdr + 24, g.btree, 8); put(g.ohdr + 32, g.heap, 8);
    // heap: names, then one free block
    bytes(g.heap, "HEAP", 4); put(g.heap + 8, g.heapSize, 8); put(g.heap + 24, g.heapData, 8);
    put(g.heapData + g.heapSize - 1, 0, 1);
    for (const auto& c : g.children) bytes(g.heapData + c.nameOff, c.name.c_str(), c.name.size() + 1);
    if (g.heapSize - g.heapUsed >= 16) { put(g.heap + 16, g.heapUsed, 8); put(g.heapData + g.heapUsed, 1, 8); put(g.heapData + g.heapUsed + 8, g.heapSize - g.heapUsed, 8); }
    else put(g.heap + 16, 1, 8);    // H5HL_FREE_NULL
    // children sorted by name over the symbol-table nodes (eight per node), one leaf B-tree node above them
    std::vector<Group::Child> cs = g.children;
    std::sort(cs.begin(), cs.end(), [](const Group::Child& a, const Group::Child& c) { return a.name < c.name; });
    bytes(g.btree, "TREE", 4); put(g.btree + 4, 0, 1); put(g.btree + 5, 0, 1); put(g.btree + 6, g.snods.size(), 2);
    put(g.btree + 8, ~0ull, 8); put(g.btree + 16, ~0ull, 8);
    put(g.btree + 544 - 1, 0, 1);
    put(g.btree + 24, 0, 8);
    for (size_t k = 0; k < g.snods.size(); k++) {
      const uint64_t sn = g.snods[k];
      const size_t lo = k * 8, hi = std::min(cs.size(), lo + 8);
      bytes(sn, "SNOD", 4); put(sn + 4, 1, 1); put(sn + 6, hi - lo, 2);
      put(sn + 328 - 1, 0, 1);
      for (size_t i = lo; i < hi; i++) {
        const uint64_t e = sn + 8 + (i - lo) * 40;
        put(e, cs[i].nameOff, 8); put(e + 8, cs[i].ohdr, 8);
        if (cs[i].isGroup) { put(e + 16, 1, 4); put(e + 24, cs[i].btree, 8); put(e + 32, cs[i].heap, 8); }
      }
      put(g.btree + 24 + 8 + k * 16, sn, 8);                       // child k
      put(g.btree + 24 + 16 + k * 16, cs[hi - 1].nameOff, 8);      // key k+1: the largest name of the node
    }
  }
  // H5Dcreate + H5Dwrite (+ H5Acreate "ftype"): version-1 object header of 16 + 256 bytes
  void datasetBody(uint64_t oh, uint64_t dataAddr, int rank, const long long* dims, bool isFloat, uint64_t nbytes, unsigned mtime, const int* ftype) {
    put(oh, 1, 1); put(oh + 2, ftype ? 7 : 6, 2); put(oh + 4, 1, 4); put(oh + 8, 256, 4);
    uint64_t p = oh + 16;
    auto hdr = [&](int type, int size, int flags) { put(p, type, 2); put(p + 2, size, 2); put(p + 4, flags, 1); p += 8; };
    // dataspace, version 1, maximum dimensions present (= the dimensions)
    hdr(0x1, 8 + 16 * rank, 0);
    put(p, 1, 1); put(p + 1, rank, 1); put(p + 2, 1, 1);
    for (int k = 0; k < rank; k++) { put(p + 8 + 8 * k, (uint64_t)dims[k], 8); put(p + 8 + 8 * rank + 8 * k, (uint64_t)dims[k], 8); }
    p += 8 + 16 * rank;
    static const unsigned char f8[24] = {0x11, 0x20, 0x3f, 0x00, 0x08, 0, 0, 0, 0, 0, 0x40, 0, 0x34, 0x0b, 0x00, 0x34, 0xff, 0x03, 0, 0, 0, 0, 0, 0};   // IEEE little-endian binary64
    static const unsigned char i4[16] = {0x10, 0x08, 0x00, 0x00, 0x04, 0, 0, 0, 0, 0, 0x20, 0, 0, 0, 0, 0};                                           // signed little-endian 32 bits
    if (isFloat) { hdr(0x3, 24, 1); bytes(p, f8, 24); p += 24; } else { hdr(0x3, 16, 1); bytes(p, i4, 16); p += 16; }
    hdr(0x5, 8, 1); put(p, 2, 1); put(p + 1, 2, 1); put(p + 2, 2, 1); put(p + 3, 1, 1); p += 8;                  // fill value, version 2: late allocation, no value
    hdr(0x8, 24, 0); put(p, 3, 1); put(p + 1, 1, 1); put(p + 2, dataAddr, 8); put(p + 10, nbytes, 8); p += 24;  // contiguous layout, version 3
    hdr(0x12, 8, 0); put(p, 1, 1); put(p + 4, mtime, 4); p += 8;                                                 // modification time, version 1
    if (ftype) {   // attribute, version 1: "ftype", int32, simple dataspace [1]
      hdr(0xc, 64, 0);
      put(p, 1, 1); put(p + 2, 6, 2); put(p + 4, 12, 2); put(p + 6, 24, 2);
      bytes(p + 8, "ftype", 6);
      bytes(p + 16, i4, 12);
      put(p + 32, 1, 1); put(p + 33, 1, 1); put(p + 34, 1, 1); put(p + 40, 1, 8); put(p + 48, 1, 8);
      put(p + 56, (uint64_t)(uint32_t)*ftype, 4);
      p += 64;
    }
    const uint64_t rest = oh + 272 - p;
    if (rest < 8) throw std::runtime_error("HDF5Io : write : object header overflow");
    hdr(0x0, (int)(rest - 8), 0);
    put(oh + 272 - 1, 0, 1);
  }
};
}  // namespace

void write_h5(const std::string& path, const H5Mesh* mesh, const std::vector<H5Field>& fields, unsigned mtime) {
  if (!mesh && fields.empty()) throw std::runtime_error("HDFIo : write : could not find anything to write");
  H5Writer w;
  size_t rootNames = (mesh ? 8 : 0) + (fields.empty() ? 0 : 16);
  H5Writer::Group root = w.makeGroup(rootNames);
  std::vector<std::pair<H5Writer::Group, std::string>> groups;
  auto writeSet = [&](H5Writer::Group& g, const std::string& name, int rank, const long long* dims, bool isFloat, const void* data, uint64_t nbytes, const int* ftype) {
    const uint64_t oh = w.allocMeta(272);            // H5Dcreate: the header, then the link in the group
    w.link(g, name, oh, false, 0, 0);
    const uint64_t addr = nbytes ? w.allocRaw(nbytes) : ~0ull;   // H5Dwrite: late allocation of the contiguous storage
    w.datasetBody(oh, addr, rank, dims, isFloat, nbytes, mtime, ftype);
    if (nbytes) w.bytes(addr, data, (size_t)nbytes);
  };
  if (mesh) {
    if (mesh->dimNodeSpace < 1 || mesh->nodesPerCell < 1) throw std::runtime_error("HDF5Io : writeMesh : the mesh has no nodes or cells");
    H5Writer::Group g = w.makeGroup(16);
    w.link(root, "Mesh", g.ohdr, true, g.btree, g.heap);
    const long long nd[2] = {(long long)(mesh->nodes.size() / mesh->dimNodeSpace), mesh->dimNodeSpace};
    const long long cd[2] = {(long long)(mesh->cells.size() / mesh->nodesPerCell), mesh->nodesPerCell};
    writeSet(g, "Nodes", 2, nd, true, mesh->nodes.data(), (uint64_t)mesh->nodes.size() * 8, nullptr);      // HDF5Io.cpp:283-301
    writeSet(g, "Cells", 2, cd, false, mesh->cells.data(), (uint64_t)mesh->cells.size() * 4, nullptr);
    w.finishGroup(g);
  }
  if (!fields.empty()) {
    size_t nb = 0;
    for (const H5Field& f : fields) nb += ((f.name.size() + 1 + 7) / 8) * 8;
    if (fields.size() > 256) throw std::runtime_error("HDF5Io : writeFields : more than 256 fields in one file are not supported");
    H5Writer::Group g = w.makeGroup(nb);
    w.link(root, "FieldData", g.ohdr, true, g.btree, g.heap);
    std::vector<const H5Field*> order;
    for (const H5Field& f : fields) order.push_back(&f);
    std::sort(order.begin(), order.end(), [](const H5Field* a, const H5Field* c) { return a->name < c->name; });   // std::map order (HDF5Io.cpp:311)
    for (const H5Field* f : order) {
      const uint64_t n = (uint64_t)f->shape[0] * (uint64_t)f->shape[1] * (uint64_t)f->shape[2];
      if (n != f->vals.size()) throw std::runtime_error("HDF5Io : writeFields : problem writing values of field: " + f->name);
      writeSet(g, f->name, 3, f->shape, true, f->vals.data(), n * 8, &f->ftype);
    }
    w.finishGroup(g);
  }
  w.finishGroup(root);
  w.close();
  // superblock, version 0 (root symbol-table entry with cached B-tree / heap addresses)
  static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
  w.bytes(0, sig, 8);
  w.put(13, 8, 1); w.put(14, 8, 1); w.put(16, 4, 2); w.put(18, 16, 2);
  w.put(24, 0, 8); w.put(32, ~0ull, 8); w.put(40, w.eoa, 8); w.put(48, ~0ull, 8);
  w.put(56, 0, 8); w.put(64, root.ohdr, 8); w.put(72, 1, 4); w.put(80, root.btree, 8); w.put(88, root.heap, 8);
  w.b.resize((size_t)w.eoa, 0);
  std::ofstream out(path, std::ios::binary | std::ios::trunc);
  if (!out) throw std::runtime_error("HDF5Io : write : could not open " + path + " for writing");
  out.write(reinterpret_cast<const char*>(w.b.data()), (std::streamsize)w.b.size());
  if (!out) throw std::runtime_error("HDF5Io : write : could not write " + path);
}

}  // namespace hfx
