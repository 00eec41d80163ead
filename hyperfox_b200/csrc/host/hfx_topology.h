#pragma once
#include <vector>
namespace hfx {
class RefElement;
struct MeshTopology {
  int nFaces = 0;
  std::vector<int> faces;      // [nFaces][nNf]
  std::vector<int> cell2face;  // [nCells][nFc]
  std::vector<int> face2cell;  // [nFaces][2]
  std::vector<int> boundary;   // ascending global face ids
};
void compute_faces(const RefElement& re, int nCells, const int* cells, MeshTopology* out);
}  // namespace hfx
