// Mesh3D — drop-in for the reference class of the same name (/root/reference/src/Mesh3d.cuh:37-177):
// .dat loader, per-cell normals/centres/areas, neighbour-classified pair lists.  All device work goes through
// the C ABI (include/i2_abi.h: i2_mesh_geometry, i2_classify_count, i2_classify_fill).
// Difference a user can observe: the three pair lists come out sorted by (i, j) and are sized exactly
// (the reference's order is atomicAdd order and its capacities are heuristic, src/Mesh3d.cu:243-262).
#ifndef MESH3D_CUH
#define MESH3D_CUH

#include "common/cuda_math.cuh"
#include "common/device_vector.cuh"

#include <array>
#include <string>
#include <vector>

enum class neighbour_type_enum {
    simple_neighbors = 0,    // exactly one shared vertex
    attached_neighbors = 1,  // a shared edge
    not_neighbors = 2,       // nothing shared
    undefined = -1
};

class Mesh3D {
public:
    virtual ~Mesh3D() = default;

    bool loadMeshFromFile(const std::string &filename, double scale = 1.0);
    void prepareMesh();

    const auto &getSimpleNeighbors() const { return pairLists[0]; }
    const auto &getAttachedNeighbors() const { return pairLists[1]; }
    // the regular list (N^2/2 entries) is materialised on first use when the evaluator runs multi-GPU (env I2_GPUS > 1), where
    // every GPU builds its own shard of it instead
    const deviceVector<int3> &getNotNeighbors() const { materialiseNotNeighbors(); return pairLists[2]; }
    // host copies of the mesh as loaded (scaled): the multi-GPU prepare uploads them to every GPU
    const std::vector<Point3> &getHostVertices() const { return hostVertices; }
    const std::vector<int3> &getHostCells() const { return hostCells; }
    long long getPairCount(neighbour_type_enum t) const { return ((int)t >= 0 && (int)t < 3) ? pairCounts[(int)t] : 0; }
    const auto &getVertices() const { return vertices; }
    const auto &getCells() const { return cells; }
    const auto &getCellNormals() const { return cellNormals; }
    const auto &getCellMeasures() const { return cellMeasures; }

private:
    deviceVector<Point3> vertices;
    deviceVector<int3> cells;
    deviceVector<Point3> cellNormals;
    deviceVector<Point3> cellCenters;
    deviceVector<double> cellMeasures;
    mutable deviceVector<int3> pairLists[3];  // indexed by neighbour_type_enum
    mutable bool notNeighborsDeferred = false;
    long long pairCounts[3] = {0, 0, 0};
    std::vector<Point3> hostVertices;
    std::vector<int3> hostCells;
    void materialiseNotNeighbors() const;
};

void exportMeshToObj(const std::string &filename, const std::vector<Point3> &vertices, const std::vector<int3> &cells);
void exportMeshToVtk(const std::string &filename, const std::vector<Point3> &vertices, const std::vector<int3> &cells,
                     const std::array<std::vector<unsigned char>, 3> &refinementsRequired);
std::string neighborTypeString(neighbour_type_enum neighborType);

#endif  // MESH3D_CUH
