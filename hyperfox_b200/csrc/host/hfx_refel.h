// Host-side reference-element table builder of the product (C++17, no dependencies).
// Produces the constant tensors the device kernels consume; same node ordering / face-node maps / cubature as the reference's
// ReferenceElement + Cubature (src/element/ReferenceElement.cpp:5-26, Cubature.cpp:52-59) so that existing meshes and
// fields keep their meaning.  Independent of oracle/ (the oracle has its own numpy restatement).
#pragma once
#include <string>
#include <vector>

namespace hfx {

enum Geometry { kSimplex = 0, kOrthotope = 1 };

struct CubatureRule {
  int dim = 0, degree = 0, nIP = 0;
  std::vector<double> coords;   // [nIP][dim]
  std::vector<double> weights;  // [nIP]
};

// Looks the rule up in the generated tables (hfx_tables_data.inc). Throws std::runtime_error("Cubature : ...").
CubatureRule make_cubature(int dim, int degree, Geometry g);

class RefElement {
 public:
  RefElement(int dim, int order, Geometry g);
  RefElement(int dim, int order, const std::string& geom);
  int dim() const { return dim_; }
  int order() const { return order_; }
  Geometry geometry() const { return geom_; }
  int numNodes() const { return nN_; }
  int numFaces() const { return nFc_; }
  int numIPs() const { return cub_.nIP; }
  const std::vector<double>& nodes() const { return nodes_; }             // [nN][dim]
  const std::vector<int>& faceNodes() const { return faceNodes_; }         // [nFc][nNf]
  const std::vector<int>& innerNodes() const { return innerNodes_; }
  const std::vector<double>& ipCoords() const { return cub_.coords; }
  const std::vector<double>& ipWeights() const { return cub_.weights; }
  const std::vector<double>& ipShape() const { return ipShape_; }          // [nIP][nN]
  const std::vector<double>& ipDShape() const { return ipDShape_; }        // [nIP][nN][dim]
  const RefElement* faceElement() const { return face_; }
  std::vector<double> interpolate(const double* pt) const;                 // [nN]
  std::vector<double> interpolateDeriv(const double* pt) const;            // [nN][dim]
  ~RefElement();
  RefElement(const RefElement&) = delete;
  RefElement& operator=(const RefElement&) = delete;

 private:
  void build();
  std::vector<double> modes(const double* pt) const;
  std::vector<double> derivModes(const double* pt) const;
  int dim_, order_, nN_ = 0, nFc_ = 0;
  Geometry geom_;
  std::vector<double> nodes_, invV_, ipShape_, ipDShape_;
  std::vector<int> faceNodes_, innerNodes_, modeMap_;
  CubatureRule cub_;
  RefElement* face_ = nullptr;
};

// Lobatto-grid node set of the (dim, order, geom) element, reference ordering (vertices, edges, faces, interior).
const std::vector<double>& node_set(int dim, int order, Geometry g);

}  // namespace hfx
