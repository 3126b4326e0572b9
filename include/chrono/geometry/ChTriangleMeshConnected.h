// Minimal stand-in for chrono::ChTriangleMeshConnected (reference: src/chrono/geometry/ChTriangleMeshConnected.h), the
// value type that leaks through ChSystemDemMesh::AddMesh (src/chrono_dem/physics/ChSystemDem.h:413).  Eigen-free:
// vertex / face storage, Wavefront OBJ reading (v / f records, polygons fanned into triangles, negative indices) and
// the affine Transform the Dem module applies after loading (ChSystemDem.cpp:509-523).
#ifndef CHRONO_B200_CHTRIANGLEMESHCONNECTED_H
#define CHRONO_B200_CHTRIANGLEMESHCONNECTED_H
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "chrono/core/ChMatrix33.h"
#include "chrono/core/ChVector3.h"

namespace chrono {

/// Triangle by its three vertices (reference: src/chrono/geometry/ChTriangle.h).
struct ChTriangle {
    ChVector3d p1, p2, p3;
};

class ChTriangleMeshConnected {
  public:
    ChTriangleMeshConnected() {}

    std::vector<ChVector3d>& GetCoordsVertices() { return m_vertices; }
    const std::vector<ChVector3d>& GetCoordsVertices() const { return m_vertices; }
    std::vector<ChVector3i>& GetIndicesVertices() { return m_face_v_indices; }
    const std::vector<ChVector3i>& GetIndicesVertices() const { return m_face_v_indices; }
    std::vector<ChVector3d>& GetCoordsNormals() { return m_normals; }

    unsigned int GetNumTriangles() const { return (unsigned int)m_face_v_indices.size(); }
    unsigned int GetNumVertices() const { return (unsigned int)m_vertices.size(); }
    ChTriangle GetTriangle(unsigned int i) const {
        const ChVector3i& f = m_face_v_indices[i];
        return ChTriangle{m_vertices[(size_t)f.x()], m_vertices[(size_t)f.y()], m_vertices[(size_t)f.z()]};
    }
    void AddTriangle(const ChVector3d& a, const ChVector3d& b, const ChVector3d& c) {
        const int base = (int)m_vertices.size();
        m_vertices.push_back(a); m_vertices.push_back(b); m_vertices.push_back(c);
        m_face_v_indices.push_back(ChVector3i(base, base + 1, base + 2));
    }
    void Clear() { m_vertices.clear(); m_normals.clear(); m_face_v_indices.clear(); }

    /// v' = rotscale * v + displ for every vertex.
    void Transform(const ChVector3d& displ, const ChMatrix33<double>& rotscale) {
        for (auto& v : m_vertices)
            v = rotscale * v + displ;
    }

    /// Reads 'v' and 'f' records of a Wavefront OBJ file; faces with more than three corners are fanned.
    bool LoadWavefrontMesh(const std::string& filename, bool load_normals = true, bool load_uv = false) {
        (void)load_normals; (void)load_uv;
        std::ifstream in(filename);
        if (!in)
            return false;
        Clear();
        std::string line;
        while (std::getline(in, line)) {
            std::istringstream ls(line);
            std::string tag;
            if (!(ls >> tag))
                continue;
            if (tag == "v") {
                double x, y, z;
                if (ls >> x >> y >> z)
                    m_vertices.push_back(ChVector3d(x, y, z));
            } else if (tag == "vn") {
                double x, y, z;
                if (ls >> x >> y >> z)
                    m_normals.push_back(ChVector3d(x, y, z));
            } else if (tag == "f") {
                std::vector<int> idx;
                std::string tok;
                while (ls >> tok) {
                    const int v = std::stoi(tok.substr(0, tok.find('/')));
                    idx.push_back(v > 0 ? v - 1 : (int)m_vertices.size() + v);
                }
                for (size_t k = 1; k + 1 < idx.size(); k++)
                    m_face_v_indices.push_back(ChVector3i(idx[0], idx[k], idx[k + 1]));
            }
        }
        for (const auto& f : m_face_v_indices)
            for (unsigned k = 0; k < 3; k++)
                if (f[k] < 0 || (size_t)f[k] >= m_vertices.size())
                    return false;
        return true;
    }

    static std::shared_ptr<ChTriangleMeshConnected> CreateFromWavefrontFile(const std::string& filename,
                                                                            bool load_normals = true, bool load_uv = false) {
        auto m = std::make_shared<ChTriangleMeshConnected>();
        if (!m->LoadWavefrontMesh(filename, load_normals, load_uv))
            return nullptr;
        return m;
    }

  private:
    std::vector<ChVector3d> m_vertices;
    std::vector<ChVector3d> m_normals;
    std::vector<ChVector3i> m_face_v_indices;
};

}  // namespace chrono
#endif
