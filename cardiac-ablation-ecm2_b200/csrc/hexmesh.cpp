// Host-side synthetic problem builder (no GPU, no MFEM): structured hex meshes numbered the way
// the reference numbers them, so that the arrays handed to the kernels are the ones an MFEM host
// would hand over (tests/test_hexmesh.py pins this against reference dumps, tests/golden/numbering.npz).
//
// Reference behaviour reproduced (not its code):
//   * Mesh::MakeCartesian3D: lattice vertex ids x + (y + z (ny+1)) (nx+1); hexes visited along a
//     generalised Hilbert curve (mesh/mesh.cpp:3683-3775, mesh/ncmesh.cpp:5435-5620).
//   * edges / faces are numbered in order of first appearance when walking elements in order and
//     local edges / faces in the reference-element order (mesh/mesh.cpp GetElementToEdgeTable,
//     GetElementToFaceTable; fem/geom.cpp:1020-1036); a face keeps the vertex order of the element
//     that created it (mesh/mesh.cpp:8519-8544).
//   * H1 dofs: vertices | edges | faces | interiors (fem/fespace.cpp:3426-3533); edge dofs run from
//     the lower to the higher vertex id; face dofs are laid out in the creating element's face
//     frame; lexicographic E-ordering (fem/restriction.cpp:44-62, fem/fe/fe_base.cpp dof map).
// Because the space is conforming, "which global dof sits at lattice point X" fully determines
// the gather map; the builder therefore assigns every dof a lattice point and looks points up,
// instead of carrying orientation tables around.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200pa.h"

namespace b200pa
{
extern thread_local std::string g_err;
}

namespace
{
int hfail(const char *m)
{
   b200pa::g_err = std::string("b200pa: ") + m;
   return 1;
}

struct I3
{
   int v[3];
   int sum() const { return v[0] + v[1] + v[2]; }
   int len() const { return std::abs(sum()); }
   I3 half() const { return {{v[0] / 2, v[1] / 2, v[2] / 2}}; }
   I3 unit() const { return {{(v[0] > 0) - (v[0] < 0), (v[1] > 0) - (v[1] < 0), (v[2] > 0) - (v[2] < 0)}}; }
};
inline I3 operator+(I3 a, I3 b) { return {{a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]}}; }
inline I3 operator-(I3 a, I3 b) { return {{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]}}; }
inline I3 operator-(I3 a) { return {{-a.v[0], -a.v[1], -a.v[2]}}; }

// Generalised Hilbert curve over a w x h x d block spanned by the axis vectors a, b, c
// (J. Cerveny's "gilbert" construction, which the reference uses for MakeCartesian3D).
void gilbert3d(I3 o, I3 a, I3 b, I3 c, std::vector<int> &out)
{
   const int w = a.len(), h = b.len(), d = c.len();
   const I3 da = a.unit(), db = b.unit(), dc = c.unit();
   auto line = [&](int n, I3 step)
   {
      for (int i = 0; i < n; ++i, o = o + step) { out.push_back(o.v[0]); out.push_back(o.v[1]); out.push_back(o.v[2]); }
   };
   if (h == 1 && d == 1) { line(w, da); return; }
   if (w == 1 && d == 1) { line(h, db); return; }
   if (w == 1 && h == 1) { line(d, dc); return; }

   I3 a2 = a.half(), b2 = b.half(), c2 = c.half();
   if ((a2.len() & 1) && w > 2) { a2 = a2 + da; }
   if ((b2.len() & 1) && h > 2) { b2 = b2 + db; }
   if ((c2.len() & 1) && d > 2) { c2 = c2 + dc; }

   if (2 * w > 3 * h && 2 * w > 3 * d) // long in a: split a only
   {
      gilbert3d(o, a2, b, c, out);
      gilbert3d(o + a2, a - a2, b, c, out);
   }
   else if (3 * h > 4 * d) // flat in c: split a and b
   {
      gilbert3d(o, b2, c, a2, out);
      gilbert3d(o + b2, a, b - b2, c, out);
      gilbert3d(o + (a - da) + (b2 - db), -b2, c, -(a - a2), out);
   }
   else if (3 * d > 4 * h) // flat in b: split a and c
   {
      gilbert3d(o, c2, a2, b, out);
      gilbert3d(o + c2, a, b, c - c2, out);
      gilbert3d(o + (a - da) + (c2 - dc), -c2, -(a - a2), b, out);
   }
   else // split all three
   {
      gilbert3d(o, b2, c2, a2, out);
      gilbert3d(o + b2, c, a2, b - b2, out);
      gilbert3d(o + (b2 - db) + (c - dc), a, -b2, -(c - c2), out);
      gilbert3d(o + (a - da) + b2 + (c - dc), -c, -(a - a2), b - b2, out);
      gilbert3d(o + (a - da) + (b2 - db), -b2, c2, -(a - a2), out);
   }
}

void sfc_order(int nx, int ny, int nz, std::vector<int> &xyz)
{
   xyz.clear();
   xyz.reserve(3 * (size_t)nx * ny * nz);
   const I3 X = {{nx, 0, 0}}, Y = {{0, ny, 0}}, Z = {{0, 0, nz}}, O = {{0, 0, 0}};
   if (nx >= ny && nx >= nz) { gilbert3d(O, X, Y, Z, xyz); }
   else if (ny >= nx && ny >= nz) { gilbert3d(O, Y, X, Z, xyz); }
   else { gilbert3d(O, Z, X, Y, xyz); }
}

// corner offsets of the reference hexahedron's 8 vertices
const int CV[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
const int HEX_EDGE[12][2] = {{0, 1}, {1, 2}, {3, 2}, {0, 3}, {4, 5}, {5, 6}, {7, 6}, {4, 7}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
const int HEX_FACE[6][4] = {{3, 2, 1, 0}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};

struct HexTopo
{
   int nx, ny, nz, p;
   long long ne, nv, nedges, nfaces, ndofs;
   std::vector<int> sfc;          // 3*ne lattice position of element k
   std::vector<int> edge_id[3];   // per axis, indexed by the lower vertex id
   std::vector<int> face_id[3];   // per normal axis, indexed by the lowest vertex id
   std::vector<unsigned char> face_frame[3]; // bits 0-1: corner of v0 in the two in-plane axes; bit 2: dir(v0->v1) is the 2nd in-plane axis
   inline long long vtx(int x, int y, int z) const { return x + ((long long)y + (long long)z * (ny + 1)) * (nx + 1); }

   void build(int nx_, int ny_, int nz_, int p_)
   {
      nx = nx_; ny = ny_; nz = nz_; p = p_;
      ne = (long long)nx * ny * nz;
      nv = (long long)(nx + 1) * (ny + 1) * (nz + 1);
      sfc_order(nx, ny, nz, sfc);
      for (int a = 0; a < 3; ++a)
      {
         edge_id[a].assign(nv, -1);
         face_id[a].assign(nv, -1);
         face_frame[a].assign(nv, 0);
      }
      // edges: first appearance over (element order, local edge order)
      nedges = 0;
      for (long long k = 0; k < ne; ++k)
      {
         const int ex = sfc[3 * k], ey = sfc[3 * k + 1], ez = sfc[3 * k + 2];
         for (int j = 0; j < 12; ++j)
         {
            const int *c0 = CV[HEX_EDGE[j][0]], *c1 = CV[HEX_EDGE[j][1]];
            const int axis = c0[0] != c1[0] ? 0 : (c0[1] != c1[1] ? 1 : 2);
            // every local edge of the reference hex points in +axis, so c0 is the lower end
            int &id = edge_id[axis][vtx(ex + c0[0], ey + c0[1], ez + c0[2])];
            if (id < 0) { id = (int)nedges++; }
         }
      }
      // faces: first appearance; the creating element's local vertex order is the face frame
      nfaces = 0;
      for (long long k = 0; k < ne; ++k)
      {
         const int ex = sfc[3 * k], ey = sfc[3 * k + 1], ez = sfc[3 * k + 2];
         for (int j = 0; j < 6; ++j)
         {
            const int *f = HEX_FACE[j];
            int lo[3] = {1, 1, 1}, hi[3] = {0, 0, 0};
            for (int m = 0; m < 4; ++m)
            {
               for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], CV[f[m]][a]); hi[a] = std::max(hi[a], CV[f[m]][a]); }
            }
            const int nrm = lo[0] == hi[0] ? 0 : (lo[1] == hi[1] ? 1 : 2);
            const int a1 = nrm == 0 ? 1 : 0, a2 = nrm == 2 ? 1 : 2; // in-plane axes, ascending
            const long long key = vtx(ex + lo[0], ey + lo[1], ez + lo[2]);
            int &id = face_id[nrm][key];
            if (id < 0)
            {
               id = (int)nfaces++;
               const int *v0 = CV[f[0]], *v1 = CV[f[1]];
               unsigned char fr = (unsigned char)((v0[a1] ? 1 : 0) | (v0[a2] ? 2 : 0));
               if (v1[a1] == v0[a1]) { fr |= 4; } // v0->v1 runs along a2
               face_frame[nrm][key] = fr;
            }
         }
      }
      const long long pm1 = p - 1;
      ndofs = nv + nedges * pm1 + nfaces * pm1 * pm1 + ne * pm1 * pm1 * pm1;
   }

   // global dof at local lattice point (i,j,k) in [0,p]^3 of element number `el` at (ex,ey,ez)
   inline int dof_at(long long el, int ex, int ey, int ez, int i, int j, int k) const
   {
      const int l[3] = {i, j, k};
      const int e[3] = {ex, ey, ez};
      int nb = 0, free_axis[3], nfree = 0;
      for (int a = 0; a < 3; ++a)
      {
         if (l[a] == 0 || l[a] == p) { ++nb; }
         else { free_axis[nfree++] = a; }
      }
      const long long pm1 = p - 1;
      if (nb == 3) { return (int)vtx(ex + (i == p), ey + (j == p), ez + (k == p)); }
      if (nb == 2)
      {
         const int a = free_axis[0];
         int c[3] = {e[0] + (l[0] == p), e[1] + (l[1] == p), e[2] + (l[2] == p)};
         c[a] = e[a];
         return (int)(nv + edge_id[a][vtx(c[0], c[1], c[2])] * pm1 + (l[a] - 1));
      }
      if (nb == 1)
      {
         const int nrm = 3 - free_axis[0] - free_axis[1];
         const int a1 = free_axis[0], a2 = free_axis[1];
         int c[3] = {e[0], e[1], e[2]};
         c[nrm] += (l[nrm] == p);
         const long long key = vtx(c[0], c[1], c[2]);
         const unsigned char fr = face_frame[nrm][key];
         // coordinates measured from the frame's v0 corner along its two directions
         const int u1 = (fr & 1) ? p - l[a1] : l[a1]; // distance from v0 along a1
         const int u2 = (fr & 2) ? p - l[a2] : l[a2];
         const int fi = (fr & 4) ? u2 : u1, fj = (fr & 4) ? u1 : u2; // i along v0->v1, j along v0->v3
         return (int)(nv + nedges * pm1 + face_id[nrm][key] * pm1 * pm1 + (fi - 1) + (fj - 1) * pm1);
      }
      return (int)(nv + nedges * pm1 + nfaces * pm1 * pm1 + el * pm1 * pm1 * pm1 + (i - 1) + (j - 1) * pm1 + (k - 1) * pm1 * pm1);
   }
};

// ---- 1-D basis ---------------------------------------------------------------------------
void legendre(int n, double x, double &P, double &dP)
{
   // P_n and P_n' at x in [-1,1] by the three-term recurrence
   double p0 = 1.0, p1 = x, d0 = 0.0, d1 = 1.0;
   if (n == 0) { P = 1.0; dP = 0.0; return; }
   for (int j = 2; j <= n; ++j)
   {
      const double pj = ((2.0 * j - 1.0) * x * p1 - (j - 1.0) * p0) / j;
      const double dj = d0 + (2.0 * j - 1.0) * p1;
      p0 = p1; p1 = pj; d0 = d1; d1 = dj;
   }
   P = p1; dP = d1;
}

void gauss_legendre(int n, double *x, double *w) // on [0,1]
{
   for (int i = 0; i < (n + 1) / 2; ++i)
   {
      double z = cos(M_PI * (i + 0.75) / (n + 0.5)), P, dP;
      for (int it = 0; it < 100; ++it)
      {
         legendre(n, z, P, dP);
         const double dz = P / dP;
         z -= dz;
         if (fabs(dz) < 1e-16) { break; }
      }
      legendre(n, z, P, dP);
      const double wt = 2.0 / ((1.0 - z * z) * dP * dP);
      x[i] = 0.5 * (1.0 - z); x[n - 1 - i] = 0.5 * (1.0 + z);
      w[i] = w[n - 1 - i] = 0.5 * wt;
   }
}

void gauss_lobatto(int np, double *x) // np = p+1 points on [0,1]
{
   const int p = np - 1;
   x[0] = 0.0; x[p] = 1.0;
   for (int i = 1; i <= p / 2; ++i)
   {
      double z = -cos(M_PI * i / p), P, dP;
      for (int it = 0; it < 100; ++it)
      {
         legendre(p, z, P, dP);
         const double d2P = (2.0 * z * dP - p * (p + 1.0) * P) / (1.0 - z * z);
         const double dz = dP / d2P;
         z -= dz;
         if (fabs(dz) < 1e-16) { break; }
      }
      x[i] = 0.5 * (1.0 + z); x[p - i] = 0.5 * (1.0 - z);
   }
   if (p % 2 == 0) { x[p / 2] = 0.5; }
}
} // namespace

extern "C" int b200pa_hex_sizes(int nx, int ny, int nz, int p, long long *ne, long long *nv, long long *ndofs)
{
   if (nx < 1 || ny < 1 || nz < 1 || p < 1) { return hfail("hex_sizes: nx, ny, nz, p must be >= 1"); }
   const long long e = (long long)nx * ny * nz, v = (long long)(nx + 1) * (ny + 1) * (nz + 1);
   if (ne) { *ne = e; }
   if (nv) { *nv = v; }
   if (ndofs) { *ndofs = ((long long)nx * p + 1) * ((long long)ny * p + 1) * ((long long)nz * p + 1); }
   return 0;
}

extern "C" int b200pa_hex_build_part(int GNX, int GNY, int GNZ, int ox, int oy, int oz, int nx, int ny, int nz, int p, double sx,
                                     double sy, double sz, int skew, int *gather_map, int *elem_vertices, double *vertices,
                                     int *elem_ijk, unsigned char *bdr_attr_of_dof, int *lattice)
{
   if (nx < 1 || ny < 1 || nz < 1 || p < 1) { return hfail("hex_build: nx, ny, nz, p must be >= 1"); }
   if (ox < 0 || oy < 0 || oz < 0 || ox + nx > GNX || oy + ny > GNY || oz + nz > GNZ) { return hfail("hex_build: part outside the global mesh"); }
   long long ndofs = 0;
   b200pa_hex_sizes(nx, ny, nz, p, nullptr, nullptr, &ndofs);
   if (ndofs >= (1LL << 31) || (long long)nx * ny * nz * (p + 1) * (p + 1) * (p + 1) >= (1LL << 31))
   {
      return hfail("hex_build: mesh too large for int32 indices (partition it over ranks)");
   }
   HexTopo T;
   T.build(nx, ny, nz, p);
   if (T.ndofs != ndofs) { return hfail("hex_build: internal error (dof count)"); }
   const int D = p + 1;
   if (vertices)
   {
      for (int z = 0; z <= nz; ++z)
         for (int y = 0; y <= ny; ++y)
            for (int x = 0; x <= nx; ++x)
            {
               double *v = vertices + 3 * T.vtx(x, y, z);
               v[0] = ((double)(ox + x) / GNX) * sx;
               v[1] = ((double)(oy + y) / GNY) * sy;
               v[2] = ((double)(oz + z) / GNZ) * sz;
               if (skew) { v[1] += 0.2 * v[0]; v[2] += 0.3 * v[0]; }
            }
   }
   const bool want_dofinfo = bdr_attr_of_dof || lattice;
   if (bdr_attr_of_dof) { std::memset(bdr_attr_of_dof, 0, (size_t)ndofs); }
   for (long long k = 0; k < T.ne; ++k)
   {
      const int ex = T.sfc[3 * k], ey = T.sfc[3 * k + 1], ez = T.sfc[3 * k + 2];
      if (elem_ijk) { elem_ijk[3 * k] = ex; elem_ijk[3 * k + 1] = ey; elem_ijk[3 * k + 2] = ez; }
      if (elem_vertices)
      {
         for (int m = 0; m < 8; ++m) { elem_vertices[8 * k + m] = (int)T.vtx(ex + CV[m][0], ey + CV[m][1], ez + CV[m][2]); }
      }
      if (!gather_map && !want_dofinfo) { continue; }
      for (int kk = 0; kk < D; ++kk)
         for (int j = 0; j < D; ++j)
            for (int i = 0; i < D; ++i)
            {
               const int g = T.dof_at(k, ex, ey, ez, i, j, kk);
               if (gather_map) { gather_map[k * D * D * D + i + D * (j + D * kk)] = g; }
               if (want_dofinfo)
               {
                  const int X = (ox + ex) * p + i, Y = (oy + ey) * p + j, Z = (oz + ez) * p + kk;
                  if (lattice) { lattice[3 * (size_t)g] = X; lattice[3 * (size_t)g + 1] = Y; lattice[3 * (size_t)g + 2] = Z; }
                  if (bdr_attr_of_dof)
                  {
                     // boundary attributes of MakeCartesian3D: z=0:1, y=0:2, x=max:3, y=max:4, x=0:5, z=max:6
                     unsigned char b = 0;
                     if (Z == 0) { b |= 1u << 0; }
                     if (Y == 0) { b |= 1u << 1; }
                     if (X == GNX * p) { b |= 1u << 2; }
                     if (Y == GNY * p) { b |= 1u << 3; }
                     if (X == 0) { b |= 1u << 4; }
                     if (Z == GNZ * p) { b |= 1u << 5; }
                     bdr_attr_of_dof[g] = b;
                  }
               }
            }
   }
   return 0;
}

extern "C" int b200pa_hex_build(int nx, int ny, int nz, int p, double sx, double sy, double sz, int skew, int *gather_map,
                                int *elem_vertices, double *vertices, int *elem_ijk, unsigned char *bdr_attr_of_dof)
{
   return b200pa_hex_build_part(nx, ny, nz, 0, 0, 0, nx, ny, nz, p, sx, sy, sz, skew, gather_map, elem_vertices, vertices, elem_ijk,
                                bdr_attr_of_dof, nullptr);
}

extern "C" int b200pa_hex_dof_lattice(int nx, int ny, int nz, int p, int *lattice)
{
   if (!lattice) { return hfail("hex_dof_lattice: NULL output"); }
   return b200pa_hex_build_part(nx, ny, nz, 0, 0, 0, nx, ny, nz, p, 1, 1, 1, 0, nullptr, nullptr, nullptr, nullptr, nullptr, lattice);
}

// ------------------------------------------------------------ results hand-off (wire formats)
// Mesh::Print in the "MFEM mesh v1.0" format (mesh/mesh.cpp:12239-12360) for the mesh b200pa_hex_build numbers:
// elements in the reference's space-filling-curve order, boundary quads and attributes exactly as
// Mesh::Make3D adds them (mesh/mesh.cpp:3806-3940), vertices with 17 significant digits.  A mesh the reference
// loads from this file gets the same H1 numbering as b200pa_hex_build (tests/test_wire_formats.py).
extern "C" int b200pa_hex_write_mesh(const char *path, int nx, int ny, int nz, double sx, double sy, double sz, int skew)
{
   if (!path || nx < 1 || ny < 1 || nz < 1) { return hfail("hex_write_mesh: bad arguments"); }
   FILE *f = std::fopen(path, "w");
   if (!f) { return hfail("hex_write_mesh: cannot open the output file"); }
   std::vector<int> sfc;
   sfc_order(nx, ny, nz, sfc);
   auto V = [&](int x, int y, int z) { return x + ((long long)y + (long long)z * (ny + 1)) * (nx + 1); };
   const long long ne = (long long)nx * ny * nz, nv = (long long)(nx + 1) * (ny + 1) * (nz + 1);
   std::fprintf(f, "MFEM mesh v1.0\n\n#\n# MFEM Geometry Types (see fem/geom.hpp):\n#\n# POINT       = 0\n# SEGMENT     = 1\n"
                   "# TRIANGLE    = 2\n# SQUARE      = 3\n# TETRAHEDRON = 4\n# CUBE        = 5\n# PRISM       = 6\n# PYRAMID     = 7\n#\n");
   std::fprintf(f, "\ndimension\n3\n\nelements\n%lld\n", ne);
   for (long long k = 0; k < ne; ++k)
   {
      const int ex = sfc[3 * k], ey = sfc[3 * k + 1], ez = sfc[3 * k + 2];
      std::fprintf(f, "1 5");
      for (int m = 0; m < 8; ++m) { std::fprintf(f, " %lld", V(ex + CV[m][0], ey + CV[m][1], ez + CV[m][2])); }
      std::fprintf(f, "\n");
   }
   std::fprintf(f, "\nboundary\n%lld\n", 2LL * ((long long)nx * ny + (long long)nx * nz + (long long)ny * nz));
   auto quad = [&](int attr, long long a, long long b, long long c, long long d) { std::fprintf(f, "%d 3 %lld %lld %lld %lld\n", attr, a, b, c, d); };
   for (int y = 0; y < ny; ++y) { for (int x = 0; x < nx; ++x) { quad(1, V(x, y, 0), V(x, y + 1, 0), V(x + 1, y + 1, 0), V(x + 1, y, 0)); } }
   for (int y = 0; y < ny; ++y) { for (int x = 0; x < nx; ++x) { quad(6, V(x, y, nz), V(x + 1, y, nz), V(x + 1, y + 1, nz), V(x, y + 1, nz)); } }
   for (int z = 0; z < nz; ++z) { for (int y = 0; y < ny; ++y) { quad(5, V(0, y, z), V(0, y, z + 1), V(0, y + 1, z + 1), V(0, y + 1, z)); } }
   for (int z = 0; z < nz; ++z) { for (int y = 0; y < ny; ++y) { quad(3, V(nx, y, z), V(nx, y + 1, z), V(nx, y + 1, z + 1), V(nx, y, z + 1)); } }
   for (int x = 0; x < nx; ++x) { for (int z = 0; z < nz; ++z) { quad(2, V(x, 0, z), V(x + 1, 0, z), V(x + 1, 0, z + 1), V(x, 0, z + 1)); } }
   for (int x = 0; x < nx; ++x) { for (int z = 0; z < nz; ++z) { quad(4, V(x, ny, z), V(x, ny, z + 1), V(x + 1, ny, z + 1), V(x + 1, ny, z)); } }
   std::fprintf(f, "\nvertices\n%lld\n3\n", nv);
   for (int z = 0; z <= nz; ++z)
      for (int y = 0; y <= ny; ++y)
         for (int x = 0; x <= nx; ++x)
         {
            double v[3] = {((double)x / nx) * sx, ((double)y / ny) * sy, ((double)z / nz) * sz};
            if (skew) { v[1] += 0.2 * v[0]; v[2] += 0.3 * v[0]; }
            std::fprintf(f, "%.17g %.17g %.17g\n", v[0], v[1], v[2]);
         }
   const bool bad = std::ferror(f) != 0;
   if (std::fclose(f) != 0 || bad) { return hfail("hex_write_mesh: write error"); }
   return 0;
}

// GridFunction::Save (fem/gridfunc.cpp:4142-4165, FiniteElementSpace::Save fem/fespace.cpp:4395-4480) of a scalar H1
// field of order p in the L-dof numbering of b200pa_hex_build: header + one value per line, 17 significant digits
extern "C" int b200pa_write_gridfunction(const char *path, int p, long long n, const double *values)
{
   if (!path || p < 1 || n < 0 || (n > 0 && !values)) { return hfail("write_gridfunction: bad arguments"); }
   FILE *f = std::fopen(path, "w");
   if (!f) { return hfail("write_gridfunction: cannot open the output file"); }
   std::fprintf(f, "FiniteElementSpace\nFiniteElementCollection: H1_3D_P%d\nVDim: 1\nOrdering: 0\n\n", p);
   for (long long i = 0; i < n; ++i) { std::fprintf(f, "%.17g\n", values[i]); }
   const bool bad = std::ferror(f) != 0;
   if (std::fclose(f) != 0 || bad) { return hfail("write_gridfunction: write error"); }
   return 0;
}

extern "C" int b200pa_randomize(int seed, long long n, double *out)
{
   if (n < 0 || (n > 0 && !out)) { return hfail("randomize: bad arguments"); }
   srand((unsigned)seed);
   const double scale = 1.0 / ((double)RAND_MAX + 1.0);
   for (long long i = 0; i < n; ++i) { out[i] = rand() * scale; }
   return 0;
}

extern "C" int b200pa_basis(int p, int q1d, double *B, double *G, double *w1d, double *W, double *gll_nodes)
{
   if (p < 1 || p > 13 || q1d < 1 || q1d > 14) { return hfail("basis: order / rule out of range"); }
   const int D = p + 1, Q = q1d;
   double xn[16], xq[16], wq[16];
   gauss_lobatto(D, xn);
   gauss_legendre(Q, xq, wq);
   if (gll_nodes) { for (int d = 0; d < D; ++d) { gll_nodes[d] = xn[d]; } }
   if (w1d) { for (int q = 0; q < Q; ++q) { w1d[q] = wq[q]; } }
   if (W)
   {
      for (int qz = 0; qz < Q; ++qz)
         for (int qy = 0; qy < Q; ++qy)
            for (int qx = 0; qx < Q; ++qx) { W[qx + Q * (qy + Q * qz)] = wq[qx] * wq[qy] * wq[qz]; }
   }
   // Lagrange basis on the GLL nodes and its derivative at the Gauss points, column-major [Q,D]
   for (int d = 0; d < D; ++d)
   {
      double den = 1.0;
      for (int m = 0; m < D; ++m) { if (m != d) { den *= (xn[d] - xn[m]); } }
      for (int q = 0; q < Q; ++q)
      {
         double val = 1.0, der = 0.0;
         for (int m = 0; m < D; ++m) { if (m != d) { val *= (xq[q] - xn[m]); } }
         for (int s = 0; s < D; ++s)
         {
            if (s == d) { continue; }
            double t = 1.0;
            for (int m = 0; m < D; ++m) { if (m != d && m != s) { t *= (xq[q] - xn[m]); } }
            der += t;
         }
         if (B) { B[q + Q * d] = val / den; }
         if (G) { G[q + Q * d] = der / den; }
      }
   }
   return 0;
}


// The 1-D matrix of the order-refinement transfer (TensorProductPRefinementTransferOperator, fem/transfer.cpp:2223-2262):
// B[q + (pf+1) d] = d-th GLL-nodal Lagrange basis function of order pc at the q-th GLL node of order pf.
extern "C" int b200pa_basis_transfer(int pc, int pf, double *B)
{
   if (pc < 1 || pf < pc || pf > 13 || !B) { return hfail("basis_transfer: orders out of range"); }
   const int DC = pc + 1, DF = pf + 1;
   double xc[16], xf[16];
   gauss_lobatto(DC, xc);
   gauss_lobatto(DF, xf);
   for (int d = 0; d < DC; ++d)
   {
      double den = 1.0;
      for (int m = 0; m < DC; ++m) { if (m != d) { den *= (xc[d] - xc[m]); } }
      for (int q = 0; q < DF; ++q)
      {
         double val = 1.0;
         for (int m = 0; m < DC; ++m) { if (m != d) { val *= (xf[q] - xc[m]); } }
         B[q + DF * d] = val / den;
      }
   }
   return 0;
}
