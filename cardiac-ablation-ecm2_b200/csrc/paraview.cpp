// Results hand-off in the reference's ParaView layout (host side, no GPU): what ParaViewDataCollection::Save writes for a
// hexahedral mesh and scalar H1 fields -
//     <prefix>/<name>/<name>.pvd                      the time series (one DataSet line per saved cycle)
//     <prefix>/<name>/Cycle000012/data.pvtu           the pieces of one cycle (one per rank)
//     <prefix>/<name>/Cycle000012/proc000003.vtu      this rank's elements
// Reference behaviour followed (not its code): fem/datacollection.cpp:887-1083 (Save: directories, PVD bookkeeping),
// :1085-1161 (PVTU header / footer, SaveDataVTU), :1163-1212 (SaveGFieldVTU); mesh/mesh.cpp:12683-12890 (Mesh::PrintVTU:
// every element carries its own (ref+1)^3 uniformly spaced points - nothing is shared between elements -, Lagrange
// hexahedra of order `ref` or ref^3 linear sub-cells, offsets, types, the element attribute as cell data);
// mesh/vtk.cpp:381-540 (VTK's node order of a Lagrange hexahedron), :560-660 (ascii, or base64 of the raw little-endian
// bytes behind a uint32 byte count); fem/geom.cpp:1315-1356 (refined reference cube).  zlib compression is not offered.
//
// The field is handed over in the L-dof numbering of the H1 space (GLL-nodal basis, the gather map of the
// ElementRestriction): element values are gathered, interpolated to the uniform points with three 1-D contractions, and
// the element's vertices give the point coordinates through the trilinear map (meshes without a nodal GridFunction).
#include <cerrno>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include <sys/stat.h>
#include <sys/types.h>

#include "../../include/b200pa.h"

namespace b200pa
{
extern thread_local std::string g_err;
}

namespace
{
int pfail(const std::string &m)
{
   b200pa::g_err = "b200pa: paraview_save: " + m;
   return 1;
}

std::string padded(long long v, int digits)
{
   char buf[64];
   snprintf(buf, sizeof(buf), "%0*lld", digits, v);
   return buf;
}

// mkdir -p (DataCollection::create_directory, fem/datacollection.cpp:35-72)
bool make_dirs(const std::string &dir)
{
   std::string::size_type pos = 0;
   do
   {
      pos = dir.find('/', pos + 1);
      const std::string sub = dir.substr(0, pos);
      if (!sub.empty() && mkdir(sub.c_str(), 0777) != 0 && errno != EEXIST) { return false; }
   }
   while (pos != std::string::npos);
   return true;
}

const char b64[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";

void base64(std::string &out, const unsigned char *in, size_t n)
{
   size_t i = 0;
   for (; i + 3 <= n; i += 3)
   {
      out += b64[in[i] >> 2];
      out += b64[((in[i] & 3) << 4) | (in[i + 1] >> 4)];
      out += b64[((in[i + 1] & 15) << 2) | (in[i + 2] >> 6)];
      out += b64[in[i + 2] & 63];
   }
   if (n - i == 1)
   {
      out += b64[in[i] >> 2];
      out += b64[(in[i] & 3) << 4];
      out += "==";
   }
   else if (n - i == 2)
   {
      out += b64[in[i] >> 2];
      out += b64[((in[i] & 3) << 4) | (in[i + 1] >> 4)];
      out += b64[(in[i + 1] & 15) << 2];
      out += '=';
   }
}

// One <DataArray> body: ascii tokens go straight to the stream, binary ones are collected and written as
// base64(uint32 byte count) base64(bytes) '\n' when the array ends
struct ArrayWriter
{
   std::ostream &os;
   int format; // 0 ascii, 1 binary (Float64), 2 binary32 (Float32)
   std::vector<unsigned char> buf;
   ArrayWriter(std::ostream &o, int f) : os(o), format(f) {}
   template <typename T> void raw(T v)
   {
      const unsigned char *p = reinterpret_cast<const unsigned char *>(&v);
      buf.insert(buf.end(), p, p + sizeof(T));
   }
   void real(double v, const char *suffix)
   {
      if (format == 2) { raw<float>((float)v); }
      else if (format == 1) { raw<double>(v); }
      else { os << ((std::fabs(v) >= DBL_MIN) ? v : 0.0) << suffix; } // ZeroSubnormal
   }
   void integer(int v, const char *suffix)
   {
      if (format == 0) { os << v << suffix; }
      else { raw<int32_t>(v); }
   }
   void byte(uint8_t v, const char *suffix)
   {
      if (format == 0) { os << (int)v << suffix; }
      else { raw<uint8_t>(v); }
   }
   void eol() { if (format == 0) { os << '\n'; } }
   bool end()
   {
      if (format != 0)
      {
         if (buf.size() > 0xffffffffULL) { return false; }
         const uint32_t nbytes = (uint32_t)buf.size();
         std::string enc;
         enc.reserve(16 + buf.size() / 3 * 4);
         base64(enc, reinterpret_cast<const unsigned char *>(&nbytes), sizeof(nbytes));
         base64(enc, buf.data(), buf.size());
         os << enc << '\n';
         buf.clear();
      }
      return true;
   }
};

// position of lattice node (i,j,k) of an order-`ref` Lagrange hexahedron in VTK's node list: corners, edges, faces,
// interior (mesh/vtk.cpp:421-489)
int vtk_hex_node(int i, int j, int k, int ref)
{
   const bool ib = (i == 0 || i == ref), jb = (j == 0 || j == ref), kb = (k == 0 || k == ref);
   const int nb = (int)ib + (int)jb + (int)kb, m = ref - 1;
   if (nb == 3) { return (i ? (j ? 2 : 1) : (j ? 3 : 0)) + (k ? 4 : 0); }
   int off = 8;
   if (nb == 2)
   {
      if (!ib) { return (i - 1) + (j ? 2 * m : 0) + (k ? 4 * m : 0) + off; }
      if (!jb) { return (j - 1) + (i ? m : 3 * m) + (k ? 4 * m : 0) + off; }
      off += 8 * m;
      return (k - 1) + m * (i ? (j ? 2 : 1) : (j ? 3 : 0)) + off;
   }
   off += 12 * m;
   if (nb == 1)
   {
      if (ib) { return (j - 1) + m * (k - 1) + (i ? m * m : 0) + off; }
      off += 2 * m * m;
      if (jb) { return (i - 1) + m * (k - 1) + (j ? m * m : 0) + off; }
      off += 2 * m * m;
      return (i - 1) + m * (j - 1) + (k ? m * m : 0) + off;
   }
   off += 6 * m * m;
   return off + (i - 1) + m * ((j - 1) + m * (k - 1));
}

} // namespace

extern "C" int b200pa_paraview_save(const char *prefix_path, const char *collection, int cycle, double time, int rank, int nranks,
                                    int p, long long ne, long long ndofs, const int *gather_map, const double *vertices,
                                    const int *elem_vertices, const int *attributes, int nfields, const char *const *names,
                                    const double *const *values_host, int levels_of_detail, int high_order, int format, int append)
{
   if (!collection || !*collection) { return pfail("no collection name"); }
   if (p < 1 || p > 13 || ne < 0 || ndofs < 0 || rank < 0 || nranks < 1 || rank >= nranks || cycle < 0) { return pfail("bad arguments"); }
   if (ne > 0 && (!gather_map || !vertices || !elem_vertices)) { return pfail("mesh arrays missing"); }
   if (nfields < 0 || (nfields > 0 && (!names || !values_host))) { return pfail("field arrays missing"); }
   if (format < 0 || format > 2) { return pfail("format: 0 ascii, 1 binary (Float64), 2 binary32 (Float32)"); }
   const int ref = levels_of_detail < 1 ? 1 : levels_of_detail;
   if (ref > 32) { return pfail("levels_of_detail out of range"); }
   const int D = p + 1, R = ref + 1, D3 = D * D * D, R3 = R * R * R;
   const long long npts = ne * (long long)R3, ncells = high_order ? ne : ne * (long long)ref * ref * ref;
   if (npts > 0x7fffffffLL) { return pfail("more than 2^31 points in one piece (VTK Int32 connectivity)"); }

   const std::string prefix = prefix_path ? prefix_path : "";
   const std::string col = prefix + collection;
   const std::string cyc = "Cycle" + padded(cycle, 6);
   if (!make_dirs(col + "/" + cyc)) { return pfail("cannot create directory " + col + "/" + cyc); }
   const char *fmt_str = format == 0 ? "ascii" : "binary";
   const char *type_str = format == 2 ? "Float32" : "Float64";

   // 1-D interpolation from the GLL nodes to the uniform points i/ref: U[i][d]
   double xn[16];
   if (b200pa_basis(p, 1, nullptr, nullptr, nullptr, nullptr, xn)) { return 1; } // the H1 basis nodes (Gauss-Lobatto on [0,1])
   std::vector<double> U((size_t)R * D);
   for (int d = 0; d < D; ++d)
   {
      double den = 1.0;
      for (int m = 0; m < D; ++m) { if (m != d) { den *= (xn[d] - xn[m]); } }
      for (int i = 0; i < R; ++i)
      {
         const double x = (double)i / ref;
         double val = 1.0;
         for (int m = 0; m < D; ++m) { if (m != d) { val *= (x - xn[m]); } }
         U[(size_t)i * D + d] = val / den;
      }
   }

   // ---- this rank's piece
   {
      const std::string path = col + "/" + cyc + "/proc" + padded(rank, 6) + ".vtu";
      std::ofstream os(path);
      if (!os.good()) { return pfail("cannot open " + path); }
      os.precision(6);
      os << "<VTKFile type=\"UnstructuredGrid\" version=\"2.2\" byte_order=\"LittleEndian\">\n";
      os << "<UnstructuredGrid>\n";
      os << "<Piece NumberOfPoints=\"" << npts << "\" NumberOfCells=\"" << ncells << "\">\n";
      ArrayWriter w(os, format);
      // points: trilinear image of the uniform lattice, x fastest
      os << "<Points>\n";
      os << "<DataArray type=\"" << type_str << "\" NumberOfComponents=\"3\" format=\"" << fmt_str << "\">\n";
      for (long long e = 0; e < ne; ++e)
      {
         const int *ev = elem_vertices + 8 * e;
         for (int k = 0; k < R; ++k)
            for (int j = 0; j < R; ++j)
               for (int i = 0; i < R; ++i)
               {
                  const double x = (double)i / ref, y = (double)j / ref, z = (double)k / ref;
                  const double ox = 1.0 - x, oy = 1.0 - y, oz = 1.0 - z;
                  const double s[8] = {ox * oy * oz, x * oy * oz, x * y * oz, ox * y * oz, ox * oy * z, x * oy * z, x * y * z, ox * y * z};
                  double pt[3] = {0.0, 0.0, 0.0};
                  for (int v = 0; v < 8; ++v)
                  {
                     const double *vx = vertices + 3 * (size_t)ev[v];
                     pt[0] += s[v] * vx[0]; pt[1] += s[v] * vx[1]; pt[2] += s[v] * vx[2];
                  }
                  w.real(pt[0], " "); w.real(pt[1], " "); w.real(pt[2], "");
                  w.eol();
               }
      }
      if (!w.end()) { return pfail("array larger than 4 GiB"); }
      os << "</DataArray>" << std::endl;
      os << "</Points>" << std::endl;

      os << "<Cells>" << std::endl;
      os << "<DataArray type=\"Int32\" Name=\"connectivity\" format=\"" << fmt_str << "\">" << std::endl;
      std::vector<int> local;
      if (high_order)
      {
         local.resize(R3);
         for (int k = 0; k < R; ++k)
            for (int j = 0; j < R; ++j)
               for (int i = 0; i < R; ++i) { local[vtk_hex_node(i, j, k, ref)] = i + R * (j + R * k); }
         for (long long e = 0; e < ne; ++e)
         {
            for (int n = 0; n < R3; ++n) { w.integer((int)(e * R3) + local[n], " "); }
            w.eol();
         }
      }
      else
      {
         for (long long e = 0; e < ne; ++e)
         {
            const int base = (int)(e * R3);
            for (int k = 0; k < ref; ++k)
               for (int j = 0; j < ref; ++j)
                  for (int i = 0; i < ref; ++i)
                  {
                     const int c[8] = {i + R * (j + R * k), i + 1 + R * (j + R * k), i + 1 + R * (j + 1 + R * k), i + R * (j + 1 + R * k),
                                       i + R * (j + R * (k + 1)), i + 1 + R * (j + R * (k + 1)), i + 1 + R * (j + 1 + R * (k + 1)),
                                       i + R * (j + 1 + R * (k + 1))};
                     for (int v = 0; v < 8; ++v) { w.integer(base + c[v], " "); }
                     w.eol();
                  }
         }
      }
      if (!w.end()) { return pfail("array larger than 4 GiB"); }
      os << "</DataArray>" << std::endl;
      os << "<DataArray type=\"Int32\" Name=\"offsets\" format=\"" << fmt_str << "\">" << std::endl;
      const int per_cell = high_order ? R3 : 8;
      for (long long c = 1; c <= ncells; ++c) { w.integer((int)(c * per_cell), "\n"); }
      if (!w.end()) { return pfail("array larger than 4 GiB"); }
      os << "</DataArray>" << std::endl;
      os << "<DataArray type=\"UInt8\" Name=\"types\" format=\"" << fmt_str << "\">" << std::endl;
      for (long long c = 0; c < ncells; ++c) { w.byte((uint8_t)(high_order ? 72 : 12), "\n"); } // VTK_LAGRANGE_HEXAHEDRON | VTK_HEXAHEDRON
      if (!w.end()) { return pfail("array larger than 4 GiB"); }
      os << "</DataArray>" << std::endl;
      os << "</Cells>" << std::endl;

      os << "<CellData Scalars=\"attribute\">" << std::endl;
      os << "<DataArray type=\"Int32\" Name=\"attribute\" format=\"" << fmt_str << "\">" << std::endl;
      const long long rep = high_order ? 1 : (long long)ref * ref * ref;
      for (long long e = 0; e < ne; ++e)
      {
         for (long long r = 0; r < rep; ++r) { w.integer(attributes ? attributes[e] : 1, "\n"); }
      }
      if (!w.end()) { return pfail("array larger than 4 GiB"); }
      os << "</DataArray>" << std::endl;
      os << "</CellData>" << std::endl;

      os << "<PointData >\n";
      std::vector<double> ev(D3), t1((size_t)R * D * D), t2((size_t)R * R * D);
      for (int f = 0; f < nfields; ++f)
      {
         if (!names[f] || !values_host[f]) { return pfail("field name / values missing"); }
         os << "<DataArray type=\"" << type_str << "\" Name=\"" << names[f] << "\" NumberOfComponents=\"1\"  format=\"" << fmt_str << "\" >" << '\n';
         const double *val = values_host[f];
         for (long long e = 0; e < ne; ++e)
         {
            const int *g = gather_map + e * D3;
            for (int n = 0; n < D3; ++n)
            {
               if (g[n] < 0 || g[n] >= ndofs) { return pfail("gather map entry out of range"); }
               ev[n] = val[g[n]];
            }
            // x, then y, then z
            for (int dz = 0; dz < D; ++dz)
               for (int dy = 0; dy < D; ++dy)
                  for (int i = 0; i < R; ++i)
                  {
                     double s = 0.0;
                     for (int dx = 0; dx < D; ++dx) { s += U[(size_t)i * D + dx] * ev[dx + D * (dy + D * dz)]; }
                     t1[i + (size_t)R * (dy + D * dz)] = s;
                  }
            for (int dz = 0; dz < D; ++dz)
               for (int j = 0; j < R; ++j)
                  for (int i = 0; i < R; ++i)
                  {
                     double s = 0.0;
                     for (int dy = 0; dy < D; ++dy) { s += U[(size_t)j * D + dy] * t1[i + (size_t)R * (dy + D * dz)]; }
                     t2[i + (size_t)R * (j + R * dz)] = s;
                  }
            for (int k = 0; k < R; ++k)
               for (int j = 0; j < R; ++j)
                  for (int i = 0; i < R; ++i)
                  {
                     double s = 0.0;
                     for (int dz = 0; dz < D; ++dz) { s += U[(size_t)k * D + dz] * t2[i + (size_t)R * (j + R * dz)]; }
                     w.real(s, "\n");
                  }
         }
         if (!w.end()) { return pfail("array larger than 4 GiB"); }
         os << "</DataArray>" << std::endl;
      }
      os << "</PointData>\n";
      os << "</Piece>\n";
      os << "</UnstructuredGrid>\n";
      os << "</VTKFile>" << std::endl;
      if (!os.good()) { return pfail("write error on " + path); }
   }
   if (rank != 0) { return 0; }

   // ---- rank 0: the cycle's PVTU and the collection's PVD
   {
      const std::string path = col + "/" + cyc + "/data.pvtu";
      std::ofstream os(path);
      if (!os.good()) { return pfail("cannot open " + path); }
      os << "<?xml version=\"1.0\"?>\n";
      os << "<VTKFile type=\"PUnstructuredGrid\" version =\"2.2\" byte_order=\"LittleEndian\">\n";
      os << "<PUnstructuredGrid GhostLevel=\"0\">\n";
      os << "<PPoints>\n";
      os << "\t<PDataArray type=\"" << type_str << "\"  Name=\"Points\" NumberOfComponents=\"3\" format=\"" << fmt_str << "\"/>\n";
      os << "</PPoints>\n";
      os << "<PCells>\n";
      os << "\t<PDataArray type=\"Int32\"  Name=\"connectivity\" NumberOfComponents=\"1\" format=\"" << fmt_str << "\"/>\n";
      os << "\t<PDataArray type=\"Int32\"  Name=\"offsets\"      NumberOfComponents=\"1\" format=\"" << fmt_str << "\"/>\n";
      os << "\t<PDataArray type=\"UInt8\"  Name=\"types\"        NumberOfComponents=\"1\" format=\"" << fmt_str << "\"/>\n";
      os << "</PCells>\n";
      os << "<PPointData>\n";
      for (int f = 0; f < nfields; ++f)
      {
         os << "<PDataArray type=\"" << type_str << "\" Name=\"" << names[f] << "\" NumberOfComponents=\"1\"  format=\"" << fmt_str << "\" />\n";
      }
      os << "</PPointData>\n";
      os << "<PCellData>\n";
      os << "\t<PDataArray type=\"Int32\" Name=\"attribute\" NumberOfComponents=\"1\" format=\"" << fmt_str << "\"/>\n";
      os << "</PCellData>\n";
      for (int r = 0; r < nranks; ++r) { os << "<Piece Source=\"proc" << padded(r, 6) << ".vtu\"/>\n"; }
      os << "</PUnstructuredGrid>\n";
      os << "</VTKFile>\n";
      if (!os.good()) { return pfail("write error on " + path); }
   }
   {
      const std::string path = col + "/" + collection + ".pvd";
      std::vector<std::string> keep;
      if (append)
      {
         std::ifstream in(path);
         std::string line;
         while (in.good() && std::getline(in, line))
         {
            if (line.find("</Collection>") != std::string::npos || line.find("</VTKFile>") != std::string::npos) { continue; }
            keep.push_back(line);
         }
      }
      std::ofstream os(path, std::ios::out | std::ios::trunc);
      if (!os.good()) { return pfail("cannot open " + path); }
      if (keep.empty())
      {
         os << "<?xml version=\"1.0\"?>\n";
         os << "<VTKFile type=\"Collection\" version=\"2.2\" byte_order=\"LittleEndian\">\n";
         os << "<Collection>\n";
      }
      else
      {
         for (const std::string &l : keep) { os << l << '\n'; }
      }
      os << "<DataSet timestep=\"" << time << "\" group=\"\" part=\"" << 0 << "\" file=\"" << cyc << "/data.pvtu\" name=\"mesh\"/>\n";
      os << "</Collection>\n";
      os << "</VTKFile>" << std::endl;
      if (!os.good()) { return pfail("write error on " + path); }
   }
   return 0;
}
