"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the reference's continuous-Galerkin path: Diffusion (src/operator/Diffusion.cpp:5-47, setDiffTensor :65-73),
Source (src/operator/Source.cpp:5-48), Convection (src/operator/Convection.cpp), Mass (src/operator/Mass.cpp), Euler (src/operator/Euler.cpp:18-37), LaplaceModel (src/model/LaplaceModel.cpp:15-52), DiffusionSource (src/model/DiffusionSource.cpp), Transport (src/model/Transport.cpp),
DirichletModel (src/model/DirichletModel.cpp:19-44) and CGSolver (src/solver/CGSolver.cpp: calcSparsityPattern :261-335, assemble :42-246, solve :248-259).
Only tests/ may import it.  Pinned by tests/test_oracle_cg.py on the reference's own known answers: TestDiffusion.cpp's monomial energies
(v^T A v = 2 p^2 / (2 p - 1) per direction), TestConvection.cpp / TestMass.cpp (w^T C u = int x^2n, 1^T M 1 = volume, x^n^T M x^n = int x^2n), TestCGSolver.cpp (lightTri, Dirichlet = 3 => Solution = 3 to 1e-12), and the regression ceilings of
tests/regression/CG/TestCGLaplace.cpp."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def element_geometry(ore, X):
    """Operator::calcJacobians / calcInvJacobians / calcDetJacobians / calcMeasure (src/operator/Operator.cpp:14-84): J[ip][r][m] = sum_i dphi_i/dxi_r x_i[m]"""
    J = np.einsum("pir,im->prm", ore.ipDShape, X)
    return np.linalg.inv(J), ore.ipWeights * np.linalg.det(J)


def diffusion_matrix(ore, X, D=None):
    """Diffusion::assemble: op(i, j) += ((D gradShapes.col(i))^T gradShapes.col(j)) dV, gradShapes.col(i) = invJ * dShape_i; D nodal values interpolated to the
    cubature points (setDiffTensor); a 1 x 1 tensor scales the identity"""
    invJ, dV = element_geometry(ore, X)
    G = np.einsum("pab,pib->pia", invJ, ore.ipDShape)
    dim = ore.dim
    if D is None:
        Dp = np.broadcast_to(np.eye(dim), (ore.nIP, dim, dim))
    else:
        D = np.asarray(D, dtype=np.float64).reshape(ore.nNodes, -1)
        if D.shape[1] == 1:
            Dp = (ore.ipShape @ D[:, 0])[:, None, None] * np.eye(dim)[None]
        else:
            Dn = D.reshape(ore.nNodes, dim, dim).transpose(0, 2, 1)          # Eigen::Map of a column-major dim x dim block
            Dp = np.einsum("pi,iab->pab", ore.ipShape, Dn)
    DG = np.einsum("pab,pib->pia", Dp, G)
    return np.einsum("p,pia,pja->ij", dV, DG, G)


def source_vector(ore, X, fIP):
    """Source::assemble (Source.cpp:24-48): F_i = sum_ip dV f(x_ip) phi_i(ip)"""
    _, dV = element_geometry(ore, X)
    return ore.ipShape.T @ (dV * fIP)


def convection_matrix(ore, X, vel):
    """Convection::setVelocity / assemble (src/operator/Convection.cpp:5-49): op(k, l) += dV (v(ip)^T invJ) . dShape_l  phi_k, v interpolated from the nodes"""
    invJ, dV = element_geometry(ore, X)
    G = np.einsum("pab,pib->pia", invJ, ore.ipDShape)
    v = ore.ipShape @ np.asarray(vel, dtype=np.float64).reshape(ore.nNodes, ore.dim)
    return np.einsum("p,pa,pla,pk->kl", dV, v, G, ore.ipShape)


def mass_matrix(ore, X):
    """Mass::assemble (src/operator/Mass.cpp:5-38): op(j, k) = sum_ip dV phi_j phi_k"""
    _, dV = element_geometry(ore, X)
    return np.einsum("p,pj,pk->jk", dV, ore.ipShape, ore.ipShape)


def euler_apply(A, F, M, uOld, dt):
    """FEModel::compute + Euler::apply, implicit (src/model/FEModel.cpp:22-33, src/operator/Euler.cpp:18-37): stiffness *= dt; rhs *= dt; stiffness += M; rhs += M u_old"""
    return dt * A + M, dt * F + M @ uOld


class CGOracle:
    """CGSolver on a whole mesh: node-based CSR with sorted columns and explicit zeros (calcSparsityPattern + PETSc AIJ), element loop (Add), then per boundary
    model zero rows + Set of the DirichletModel's identity / Dirichlet values (face-node order)."""

    def __init__(self, ore, nodes, cells, faces, boundary, diff=None, source=None, vel=None, diffusion=True, dt=0.0, solOld=None):
        """diffusion=False + vel: Transport (src/model/Transport.cpp: Convection only); dt > 0: implicit Euler with the nodal Solution solOld as the old state"""
        self.ore, self.nodes, self.cells, self.faces, self.boundary = ore, nodes, cells, faces, np.asarray(boundary)
        self.diff, self.source, self.vel, self.diffusion, self.dt, self.solOld = diff, source, vel, diffusion, dt, solOld
        self.n = nodes.shape[0]

    def pattern(self):
        rows = [set() for _ in range(self.n)]
        for c in self.cells:
            for i in c:
                rows[i].update(c.tolist())
        self.rowptr = np.zeros(self.n + 1, dtype=np.int64)
        cols = []
        for i, r in enumerate(rows):
            s = sorted(r)
            cols += s
            self.rowptr[i + 1] = self.rowptr[i] + len(s)
        self.colidx = np.array(cols, dtype=np.int32)
        return self.rowptr, self.colidx

    def assemble(self, dirichlet):
        if not hasattr(self, "rowptr"):
            self.pattern()
        A = sp.lil_matrix((self.n, self.n))
        b = np.zeros(self.n)
        xipAll = np.einsum("pi,cid->cpd", self.ore.ipShape, self.nodes[self.cells])
        for e, c in enumerate(self.cells):
            X = self.nodes[c]
            Ae = diffusion_matrix(self.ore, X, None if self.diff is None else self.diff[c]) if self.diffusion else np.zeros((c.size, c.size))
            if self.vel is not None:
                Ae = Ae + convection_matrix(self.ore, X, self.vel[c])
            Fe = source_vector(self.ore, X, np.array([self.source(p) for p in xipAll[e]])) if self.source is not None else np.zeros(c.size)
            if self.dt > 0.0:
                Ae, Fe = euler_apply(Ae, Fe, mass_matrix(self.ore, X), self.solOld[c], self.dt)
            A[np.ix_(c, c)] = A[np.ix_(c, c)].toarray() + Ae
            b[c] += Fe
        A = A.tocsr()
        bn = np.unique(self.faces[self.boundary])
        mask = np.zeros(self.n, dtype=bool); mask[bn] = True
        A = sp.diags((~mask).astype(float)) @ A + sp.diags(mask.astype(float))      # zeroOutRows + Set identity
        for F in self.boundary:
            b[self.faces[F]] = dirichlet[F]                                          # Set of the Dirichlet values, face-node order
        self.A, self.b = A.tocsr(), b
        # values on the full pattern (explicit zeros kept)
        P = sp.csr_matrix((np.ones(self.colidx.size), self.colidx, self.rowptr), shape=(self.n, self.n))
        full = (self.A + 0.0 * P).tocsr(); full.sort_indices()
        dense_vals = np.zeros(self.colidx.size)
        lut = {}
        for i in range(self.n):
            lo, hi = self.rowptr[i], self.rowptr[i + 1]
            row = dict(zip(full.indices[full.indptr[i]:full.indptr[i + 1]].tolist(), full.data[full.indptr[i]:full.indptr[i + 1]].tolist()))
            dense_vals[lo:hi] = [row.get(int(c), 0.0) for c in self.colidx[lo:hi]]
        self.vals = dense_vals
        return self.vals, self.b

    def solve(self):
        self.sol = spla.spsolve(self.A.tocsc(), self.b)
        return self.sol
