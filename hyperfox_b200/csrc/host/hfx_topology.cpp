// Host-side mesh topology builder of the product: faces, cell2Face, face2Cell, boundary set from cell connectivity.
// Data contract = reference src/mesh/Mesh.cpp:183-274,377-537 (what MOAB hands back to Mesh::computeFaces):
//   * faces are numbered by first appearance while walking cells in ascending id and local faces in reference order,
//   * face2Cell lists adjacent cells ascending (second = -1 on the boundary),
//   * a face's node list is read from its lowest-id cell through the order-p face-node map.
#include <algorithm>
#include <array>
#include <cstdint>
#include <numeric>
#include <stdexcept>
#include <vector>

#include "hfx_refel.h"
#include "hfx_topology.h"

namespace hfx {

void compute_faces(const RefElement& re, int nCells, const int* cells, MeshTopology* out) {
  const int nN = re.numNodes(), nFc = re.numFaces(), nNf = re.faceElement()->numNodes();
  RefElement skel(re.dim(), re.order() != 0 ? 1 : 0, re.geometry());
  const int nVf = skel.faceElement()->numNodes();  // vertices per face of the linear skeleton
  const std::vector<int>& sfn = skel.faceNodes();
  if (nVf > 4) throw std::runtime_error("Mesh : computeFaces : unsupported face type");
  struct Ent { std::array<int, 4> key; int64_t idx; };
  const int64_t N = (int64_t)nCells * nFc;
  std::vector<Ent> ents((size_t)N);
  for (int c = 0; c < nCells; c++)
    for (int f = 0; f < nFc; f++) {
      Ent& e = ents[(size_t)c * nFc + f];
      e.key = {-1, -1, -1, -1};
      for (int k = 0; k < nVf; k++) e.key[k] = cells[(size_t)c * nN + sfn[(size_t)f * nVf + k]];
      std::sort(e.key.begin(), e.key.begin() + nVf);
      e.idx = (int64_t)c * nFc + f;
    }
  std::sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b) { return a.key != b.key ? a.key < b.key : a.idx < b.idx; });
  // groups of identical vertex sets; the group's first entry is its first appearance
  std::vector<int64_t> firsts;
  std::vector<int64_t> groupOf((size_t)N);
  for (int64_t i = 0; i < N;) {
    int64_t j = i;
    while (j < N && ents[j].key == ents[i].key) j++;
    if (j - i > 2) throw std::runtime_error("Mesh : computeFaces : a face is shared by more than two cells");
    for (int64_t k = i; k < j; k++) groupOf[k] = (int64_t)firsts.size();
    firsts.push_back(ents[i].idx);
    i = j;
  }
  const int nFaces = (int)firsts.size();
  std::vector<int> order(nFaces), rank(nFaces);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return firsts[a] < firsts[b]; });
  for (int i = 0; i < nFaces; i++) rank[order[i]] = i;
  out->nFaces = nFaces;
  out->cell2face.assign((size_t)N, -1);
  out->face2cell.assign((size_t)nFaces * 2, -1);
  for (int64_t i = 0; i < N; i++) {
    int fid = rank[groupOf[i]];
    out->cell2face[ents[i].idx] = fid;
    int c = (int)(ents[i].idx / nFc);
    if (out->face2cell[(size_t)fid * 2] < 0) out->face2cell[(size_t)fid * 2] = c;   // entries of a group are sorted by idx => lower cell first
    else out->face2cell[(size_t)fid * 2 + 1] = c;
  }
  out->faces.resize((size_t)nFaces * nNf);
  out->boundary.clear();
  const std::vector<int>& fn = re.faceNodes();
  for (int F = 0; F < nFaces; F++) {
    int c0 = out->face2cell[(size_t)F * 2];
    int lf = -1;
    for (int k = 0; k < nFc; k++) if (out->cell2face[(size_t)c0 * nFc + k] == F) { lf = k; break; }
    for (int j = 0; j < nNf; j++) out->faces[(size_t)F * nNf + j] = cells[(size_t)c0 * nN + fn[(size_t)lf * nNf + j]];
    if (out->face2cell[(size_t)F * 2 + 1] < 0) out->boundary.push_back(F);
  }
}

}  // namespace hfx
