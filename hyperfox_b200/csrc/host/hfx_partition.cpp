// Element partition + trace-halo plan (see hfx_partition.h).  Host C++, once per mesh; deterministic: every rank computes the same global
// picture and extracts its part, so no communication is needed to set the exchange up (the reference negotiates the same lists over MPI:
// src/parallel/Partitioner.cpp:42-107, :565-826).
#include "hfx_partition.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <numeric>
#include <stdexcept>

#include "hfx_refel.h"
#include "hfx_topology.h"

namespace hfx {

void rcb_partition(int dim, long long nVerts, const double* verts, long long nCells, int nv, const int* cells, int world, int* part) {
  if (world < 1) throw std::runtime_error("Partitioner : computePartition : the number of partitions must be positive");
  if (nCells < world) throw std::runtime_error("Partitioner : computePartition : every rank must own at least one cell");
  std::vector<double> cen((size_t)nCells * dim);
  for (long long c = 0; c < nCells; c++)
    for (int d = 0; d < dim; d++) {
      double s = 0.0;
      for (int k = 0; k < nv; k++) {
        const long long v = cells[(size_t)c * nv + k];
        if (v < 0 || v >= nVerts) throw std::runtime_error("Partitioner : computePartition : cell vertex id out of range");
        s += verts[(size_t)v * dim + d];
      }
      cen[(size_t)c * dim + d] = s / nv;
    }
  struct Job { std::vector<long long> ids; int r0, k; };
  std::vector<Job> stack;
  { Job j; j.ids.resize((size_t)nCells); std::iota(j.ids.begin(), j.ids.end(), 0LL); j.r0 = 0; j.k = world; stack.push_back(std::move(j)); }
  while (!stack.empty()) {
    Job job = std::move(stack.back());
    stack.pop_back();
    if (job.k == 1) { for (long long id : job.ids) part[id] = job.r0; continue; }
    // cut perpendicular to the longest axis of the bounding box of the centroids, floor(k/2) : ceil(k/2) by cell count, ties by cell id
    int ax = 0; double best = -1.0;
    for (int d = 0; d < dim; d++) {
      double lo = 1e300, hi = -1e300;
      for (long long id : job.ids) { const double x = cen[(size_t)id * dim + d]; lo = std::min(lo, x); hi = std::max(hi, x); }
      if (hi - lo > best) { best = hi - lo; ax = d; }
    }
    const int kl = job.k / 2;
    const long long n = (long long)job.ids.size(), nl = (n * kl + job.k / 2) / job.k;
    std::sort(job.ids.begin(), job.ids.end(), [&](long long a, long long b) {
      const double xa = cen[(size_t)a * dim + ax], xb = cen[(size_t)b * dim + ax];
      return xa != xb ? xa < xb : a < b;
    });
    Job L, R;
    L.ids.assign(job.ids.begin(), job.ids.begin() + nl); R.ids.assign(job.ids.begin() + nl, job.ids.end());
    std::sort(L.ids.begin(), L.ids.end()); std::sort(R.ids.begin(), R.ids.end());
    L.r0 = job.r0; L.k = kl; R.r0 = job.r0 + kl; R.k = job.k - kl;
    stack.push_back(std::move(L)); stack.push_back(std::move(R));
  }
}

void graph_partition(int dim, int geom, long long nCells, const int* cells, int world, int* part) {
  if (world < 1) throw std::runtime_error("Partitioner : computePartition : the number of partitions must be positive");
  if (nCells < world) throw std::runtime_error("Partitioner : computePartition : every rank must own at least one cell");
  RefElement lin(dim, 1, geom == 0 ? kSimplex : kOrthotope);
  const int nFc = lin.numFaces();
  MeshTopology G;
  compute_faces(lin, (int)nCells, cells, &G);
  // dual graph in CSR form: neighbours of a cell in ascending local face order
  std::vector<int> adj((size_t)nCells * nFc, -1);
  for (long long c = 0; c < nCells; c++)
    for (int f = 0; f < nFc; f++) {
      const int F = G.cell2face[(size_t)c * nFc + f];
      const int a = G.face2cell[(size_t)F * 2], b = G.face2cell[(size_t)F * 2 + 1];
      adj[(size_t)c * nFc + f] = a == c ? b : a;
    }
  std::vector<int> mark((size_t)nCells, -1);          // id of the job a cell currently belongs to
  std::vector<int> level((size_t)nCells, 0);
  struct Job { std::vector<int> ids; int r0, k; };
  std::vector<Job> stack;
  { Job j; j.ids.resize((size_t)nCells); std::iota(j.ids.begin(), j.ids.end(), 0); j.r0 = 0; j.k = world; stack.push_back(std::move(j)); }
  int jobId = 0;
  std::vector<int> order, queue;
  while (!stack.empty()) {
    Job job = std::move(stack.back());
    stack.pop_back();
    if (job.k == 1) { for (int id : job.ids) part[id] = job.r0; continue; }
    const int me = jobId++;
    for (int id : job.ids) mark[(size_t)id] = me;
    // breadth-first order of the subset from `start` (components that are not reached are appended from their lowest remaining id)
    auto bfs = [&](int start) {
      order.clear();
      const int visited = -2 - me;                     // (distinct from every job id)
      auto run = [&](int s0) {
        queue.assign(1, s0); mark[(size_t)s0] = visited; level[(size_t)s0] = 0;
        for (size_t h = 0; h < queue.size(); h++) {
          const int c = queue[h];
          order.push_back(c);
          for (int f = 0; f < nFc; f++) {
            const int nb = adj[(size_t)c * nFc + f];
            if (nb >= 0 && mark[(size_t)nb] == me) { mark[(size_t)nb] = visited; level[(size_t)nb] = level[(size_t)c] + 1; queue.push_back(nb); }
          }
        }
      };
      run(start);
      for (int id : job.ids) if (mark[(size_t)id] == me) run(id);
      for (int id : job.ids) mark[(size_t)id] = me;    // restore for the next sweep
    };
    bfs(job.ids.front());
    const int far = order.back();                      // a cell of the last shell: pseudo-peripheral start
    bfs(far);
    const int kl = job.k / 2;
    const long long n = (long long)job.ids.size(), nl = (n * kl + job.k / 2) / job.k;
    Job L, R;
    L.ids.assign(order.begin(), order.begin() + nl); R.ids.assign(order.begin() + nl, order.end());
    std::sort(L.ids.begin(), L.ids.end()); std::sort(R.ids.begin(), R.ids.end());
    L.r0 = job.r0; L.k = kl; R.r0 = job.r0 + kl; R.k = job.k - kl;
    stack.push_back(std::move(L)); stack.push_back(std::move(R));
  }
}

void build_partition_plan(int dim, int geom, long long nCells, const int* cells, const int* part, int rank, int world, PartitionPlan* P) {
  RefElement lin(dim, 1, geom == 0 ? kSimplex : kOrthotope);
  const int nv = lin.numNodes(), nFc = lin.numFaces();
  if (nCells > 2000000000LL / nFc) throw std::runtime_error("Partitioner : update : too many cells for 32-bit face ids");
  for (long long c = 0; c < nCells; c++) if (part[c] < 0 || part[c] >= world) throw std::runtime_error("Partitioner : computePartition : rank ids must lie in [0, nPartitions)");
  MeshTopology G;
  compute_faces(lin, (int)nCells, cells, &G);
  const int nF = G.nFaces;
  auto owner = [&](int F) { return part[G.face2cell[(size_t)F * 2]]; };   // a face travels with its first adjacent cell (ZoltanPartitioner.cpp:83-133)
  P->dim = dim; P->geom = geom; P->rank = rank; P->world = world; P->nv = nv; P->nFc = nFc;
  // owned cells, ghost cells across the faces this rank owns (overlap 1: the owner of a face recomputes the element on its other side)
  std::vector<long long> owned, ghosts;
  for (long long c = 0; c < nCells; c++) if (part[c] == rank) owned.push_back(c);
  if (owned.empty()) throw std::runtime_error("Partitioner : computePartition : every rank must own at least one cell");
  // ghostOf[c]: ranks that hold cell c as a ghost (CSR)
  std::vector<int> gcount((size_t)nCells + 1, 0);
  for (int F = 0; F < nF; F++) { const int c1 = G.face2cell[(size_t)F * 2 + 1]; if (c1 >= 0 && part[c1] != owner(F)) gcount[(size_t)c1 + 1]++; }
  for (long long c = 0; c < nCells; c++) gcount[(size_t)c + 1] += gcount[(size_t)c];
  std::vector<int> gr((size_t)gcount[(size_t)nCells]), gfill(gcount.begin(), gcount.end() - 1);
  for (int F = 0; F < nF; F++) {
    const int c1 = G.face2cell[(size_t)F * 2 + 1];
    if (c1 >= 0 && part[c1] != owner(F)) { gr[(size_t)gfill[(size_t)c1]++] = owner(F); if (owner(F) == rank) ghosts.push_back(c1); }
  }
  std::sort(ghosts.begin(), ghosts.end());
  ghosts.erase(std::unique(ghosts.begin(), ghosts.end()), ghosts.end());
  P->nOwned = (long long)owned.size(); P->nGhost = (long long)ghosts.size();
  P->cellsGlobal = owned; P->cellsGlobal.insert(P->cellsGlobal.end(), ghosts.begin(), ghosts.end());
  const long long nL = (long long)P->cellsGlobal.size();
  // local vertex numbering: ascending global id
  std::vector<long long> used; used.reserve((size_t)nL * nv);
  for (long long c : P->cellsGlobal) for (int k = 0; k < nv; k++) used.push_back(cells[(size_t)c * nv + k]);
  std::sort(used.begin(), used.end()); used.erase(std::unique(used.begin(), used.end()), used.end());
  P->vertexIds = used;
  P->localCells.resize((size_t)nL * nv);
  for (long long i = 0; i < nL; i++)
    for (int k = 0; k < nv; k++)
      P->localCells[(size_t)i * nv + k] = (int)(std::lower_bound(used.begin(), used.end(), (long long)cells[(size_t)P->cellsGlobal[(size_t)i] * nv + k]) - used.begin());
  // local face numbering (same walk as the global one, over the local cells) and its map to global face ids
  MeshTopology Lt;
  compute_faces(lin, (int)nL, P->localCells.data(), &Lt);
  const int nFl = Lt.nFaces;
  P->localCell2Face = Lt.cell2face; P->localFace2Cell = Lt.face2cell;
  P->faceGlobal.assign((size_t)nFl, -1);
  for (long long i = 0; i < nL; i++)
    for (int f = 0; f < nFc; f++) P->faceGlobal[(size_t)Lt.cell2face[(size_t)i * nFc + f]] = G.cell2face[(size_t)P->cellsGlobal[(size_t)i] * nFc + f];
  P->faceOwner.resize((size_t)nFl); P->ownedFace.resize((size_t)nFl);
  for (int f = 0; f < nFl; f++) { P->faceOwner[(size_t)f] = owner((int)P->faceGlobal[(size_t)f]); P->ownedFace[(size_t)f] = P->faceOwner[(size_t)f] == rank ? 1 : 0; }
  // global -> local face id
  std::vector<std::pair<long long, int>> g2l((size_t)nFl);
  for (int f = 0; f < nFl; f++) g2l[(size_t)f] = {P->faceGlobal[(size_t)f], f};
  std::sort(g2l.begin(), g2l.end());
  auto toLocal = [&](long long g) {
    auto it = std::lower_bound(g2l.begin(), g2l.end(), std::make_pair(g, -1));
    if (it == g2l.end() || it->first != g) throw std::runtime_error("Partitioner : updateSharedInformation : a shared face is not part of the local mesh");
    return it->second;
  };
  // who holds which face: rank r holds face F if one of F's cells is owned by r or is a ghost of r.  A holder that is not the owner needs the
  // owner's trace on F: the owner sends, the holder receives.  Lists in ascending global face id.
  std::vector<std::vector<int>> sendTo((size_t)world), recvFrom((size_t)world);
  std::vector<int> holders;
  for (int F = 0; F < nF; F++) {
    const int own = owner(F);
    holders.clear();
    for (int s = 0; s < 2; s++) {
      const int c = G.face2cell[(size_t)F * 2 + s];
      if (c < 0) continue;
      holders.push_back(part[c]);
      for (int k = gcount[(size_t)c]; k < gcount[(size_t)c + 1]; k++) holders.push_back(gr[(size_t)k]);
    }
    std::sort(holders.begin(), holders.end());
    holders.erase(std::unique(holders.begin(), holders.end()), holders.end());
    for (int r : holders) {
      if (r == own) continue;
      if (own == rank) sendTo[(size_t)r].push_back(toLocal(F));
      else if (r == rank) recvFrom[(size_t)own].push_back(toLocal(F));
    }
  }
  P->nbrs.clear(); P->sendCount.clear(); P->recvCount.clear(); P->sendFaces.clear(); P->recvFaces.clear();
  for (int r = 0; r < world; r++) {
    if (r == rank || (sendTo[(size_t)r].empty() && recvFrom[(size_t)r].empty())) continue;
    P->nbrs.push_back(r);
    P->sendCount.push_back((int)sendTo[(size_t)r].size()); P->recvCount.push_back((int)recvFrom[(size_t)r].size());
    P->sendFaces.insert(P->sendFaces.end(), sendTo[(size_t)r].begin(), sendTo[(size_t)r].end());
    P->recvFaces.insert(P->recvFaces.end(), recvFrom[(size_t)r].begin(), recvFrom[(size_t)r].end());
  }
  // sharedFaceList of the reference (Partitioner.h:223): faces between a cell of this partition and a cell of another one
  P->sharedFaceList.clear();
  for (int F = 0; F < nF; F++) {
    const int c0 = G.face2cell[(size_t)F * 2], c1 = G.face2cell[(size_t)F * 2 + 1];
    if (c1 < 0 || part[c0] == part[c1]) continue;
    if (part[c0] == rank) { P->sharedFaceList.push_back(F); P->sharedFaceList.push_back(part[c1]); P->sharedFaceList.push_back(c1); }
    else if (part[c1] == rank) { P->sharedFaceList.push_back(F); P->sharedFaceList.push_back(part[c0]); P->sharedFaceList.push_back(c0); }
  }
}

void face_canonical_positions(int dim, int order, long long nFaces, int nNf, const int* faces, const long long* gv, uint8_t* canon) {
  const int nvf = dim;   // vertices of a simplex face
  if (nvf > 3) throw std::runtime_error("Partitioner : update : unsupported face type");
  if (order == 1 || dim == 1) {
    for (long long F = 0; F < nFaces; F++)
      for (int a = 0; a < nNf; a++) {
        int rk = 0;
        for (int b = 0; b < nNf; b++) if (gv[faces[(size_t)F * nNf + b]] < gv[faces[(size_t)F * nNf + a]]) rk++;
        canon[(size_t)F * nNf + a] = (uint8_t)rk;
      }
    return;
  }
  RefElement fe(dim - 1, order, kSimplex);
  if (fe.numNodes() != nNf) throw std::runtime_error("Partitioner : update : the face connectivity does not match the face element");
  const std::vector<double>& ref = fe.nodes();   // [nNf][dim-1] on [-1,1]^(dim-1)
  std::vector<double> lam((size_t)nNf * nvf);
  for (int a = 0; a < nNf; a++) {
    double s0 = 1.0;
    for (int d = 0; d < dim - 1; d++) { const double l = 0.5 * (ref[(size_t)a * (dim - 1) + d] + 1.0); lam[(size_t)a * nvf + d + 1] = l; s0 -= l; }
    lam[(size_t)a * nvf] = s0;
  }
  // one table per permutation rho (rho[k] = rank of local vertex k in ascending global id): position of local node a in the canonical order
  std::vector<int> perm(nvf); std::iota(perm.begin(), perm.end(), 0);
  std::vector<std::vector<int>> tables; std::vector<std::vector<int>> keys;
  do {
    std::vector<int> pos((size_t)nNf, -1);
    for (int a = 0; a < nNf; a++) {
      double lc[3] = {0, 0, 0};
      for (int k = 0; k < nvf; k++) lc[perm[k]] = lam[(size_t)a * nvf + k];
      int bestB = -1; double bestD = 1e300;
      for (int b = 0; b < nNf; b++) {
        double d = 0.0;
        for (int k = 0; k < nvf; k++) d = std::max(d, std::fabs(lc[k] - lam[(size_t)b * nvf + k]));
        if (d < bestD) { bestD = d; bestB = b; }
      }
      if (bestD > 1e-10) throw std::runtime_error("Partitioner : update : the face node set is not symmetric under vertex permutations");
      pos[(size_t)a] = bestB;
    }
    tables.push_back(pos); keys.push_back(perm);
  } while (std::next_permutation(perm.begin(), perm.end()));
  for (long long F = 0; F < nFaces; F++) {
    int rho[3] = {0, 0, 0};
    for (int k = 0; k < nvf; k++) {
      const long long g = gv[faces[(size_t)F * nNf + k]];
      if (g < 0) throw std::runtime_error("Partitioner : update : a face vertex has no global vertex id");
      int rk = 0;
      for (int j = 0; j < nvf; j++) if (gv[faces[(size_t)F * nNf + j]] < g) rk++;
      rho[k] = rk;
    }
    size_t ti = 0;
    for (; ti < keys.size(); ti++) { bool eq = true; for (int k = 0; k < nvf; k++) eq = eq && keys[ti][(size_t)k] == rho[k]; if (eq) break; }
    if (ti == keys.size()) throw std::runtime_error("Partitioner : update : repeated global vertex id on a face");
    for (int a = 0; a < nNf; a++) canon[(size_t)F * nNf + a] = (uint8_t)tables[ti][(size_t)a];
  }
}

void face_canonical_positions_geom(int dim, int order, int geom, long long nFaces, int nNf, const int* faces, const long long* gv, uint8_t* canon) {
  if (geom == 0 || dim <= 2) { face_canonical_positions(dim, order, nFaces, nNf, faces, gv, canon); return; }   // (a line has the same two symmetries either way)
  if (dim != 3) throw std::runtime_error("Partitioner : update : unsupported face type");
  RefElement fe(2, order, kOrthotope);
  if (fe.numNodes() != nNf) throw std::runtime_error("Partitioner : update : the face connectivity does not match the face element");
  const std::vector<double>& ref = fe.nodes();   // [nNf][2] on [-1,1]^2; the four corners come first
  double cx[4][2];
  for (int k = 0; k < 4; k++) for (int d = 0; d < 2; d++) {
    cx[k][d] = ref[(size_t)k * 2 + d];
    if (std::fabs(std::fabs(cx[k][d]) - 1.0) > 1e-12) throw std::runtime_error("Partitioner : update : the corners of the face element do not come first");
  }
  auto adjacent = [&](int a, int b) { return (cx[a][0] == cx[b][0]) != (cx[a][1] == cx[b][1]); };
  // tables[A][s]: A = corner with the smallest global id, s = 0 / 1: the lower-id neighbour is the first / second neighbour of A (in corner order)
  std::vector<int> tables((size_t)4 * 2 * nNf, -1);
  for (int A = 0; A < 4; A++) {
    int nb[2], nn = 0, D = -1;
    for (int k = 0; k < 4; k++) { if (k == A) continue; if (adjacent(A, k)) nb[nn++] = k; else D = k; }
    if (nn != 2 || D < 0) throw std::runtime_error("Partitioner : update : unexpected corner layout of the face element");
    for (int s2 = 0; s2 < 2; s2++) {
      const int B = nb[s2], C = nb[1 - s2];
      double w[4][2];
      w[A][0] = -1; w[A][1] = -1; w[B][0] = 1; w[B][1] = -1; w[C][0] = -1; w[C][1] = 1; w[D][0] = 1; w[D][1] = 1;
      for (int a = 0; a < nNf; a++) {
        double eta[2] = {0, 0};
        for (int k = 0; k < 4; k++) {
          const double N = 0.25 * (1.0 + ref[(size_t)a * 2] * cx[k][0]) * (1.0 + ref[(size_t)a * 2 + 1] * cx[k][1]);
          eta[0] += N * w[k][0]; eta[1] += N * w[k][1];
        }
        int best = -1; double bd = 1e300;
        for (int b = 0; b < nNf; b++) { const double d = std::max(std::fabs(ref[(size_t)b * 2] - eta[0]), std::fabs(ref[(size_t)b * 2 + 1] - eta[1])); if (d < bd) { bd = d; best = b; } }
        if (bd > 1e-10) throw std::runtime_error("Partitioner : update : the face node set is not symmetric under the symmetries of the square");
        tables[((size_t)A * 2 + s2) * nNf + a] = best;
      }
    }
  }
  for (long long F = 0; F < nFaces; F++) {
    long long g[4];
    for (int k = 0; k < 4; k++) { g[k] = gv[faces[(size_t)F * nNf + k]]; if (g[k] < 0) throw std::runtime_error("Partitioner : update : a face vertex has no global vertex id"); }
    int A = 0;
    for (int k = 1; k < 4; k++) if (g[k] < g[A]) A = k;
    int nb[2], nn = 0;
    for (int k = 0; k < 4; k++) if (k != A && adjacent(A, k)) nb[nn++] = k;
    const int s2 = g[nb[0]] < g[nb[1]] ? 0 : 1;
    if (g[nb[0]] == g[nb[1]]) throw std::runtime_error("Partitioner : update : repeated global vertex id on a face");
    for (int a = 0; a < nNf; a++) canon[(size_t)F * nNf + a] = (uint8_t)tables[((size_t)A * 2 + s2) * nNf + a];
  }
}

}  // namespace hfx
