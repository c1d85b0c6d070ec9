// Mesh3D host class: file parsing and exports on the host, geometry + classification through the C ABI.
#include "../../include/integrator2/Mesh3d.cuh"
#include "host_context.h"

#include <fstream>

bool Mesh3D::loadMeshFromFile(const std::string &filename, double scale) {
    std::ifstream in(filename);
    if (!in.is_open()) {
        printf("Error while opening the file\n");
        return false;
    }
    int numVertices = 0, numEntities = 0;
    in >> numVertices >> numEntities;

    hostVertices.clear();
    hostCells.clear();
    hostVertices.reserve(numVertices);
    hostCells.reserve(numEntities);
    for (int v = 0; v < numVertices; ++v) {
        int id;
        Point3 p;
        in >> id >> p.x >> p.y >> p.z;
        hostVertices.push_back(scale * p);
    }
    // entities: "id 203 a b c" is a triangle (1-based vertex ids); any other type is "id type a b" and is skipped
    int id, type;
    while (in >> id >> type) {
        if (type == 203) {
            int3 t;
            if (!(in >> t.x >> t.y >> t.z)) break;
            hostCells.push_back(make_int3(t.x - 1, t.y - 1, t.z - 1));
        } else {
            int a, b;
            in >> a >> b;
        }
    }
    const int numCells = (int)hostCells.size();

    vertices.allocate(numVertices);
    cells.allocate(numCells);
    cellNormals.allocate(numCells);
    cellCenters.allocate(numCells);
    cellMeasures.allocate(numCells);
    copy_h2d(hostVertices.data(), vertices.data, vertices.size);
    copy_h2d(hostCells.data(), cells.data, cells.size);
    checkCudaErrors(cudaDeviceSynchronize());

    printf("Loaded mesh with %d vertices and %d cells\n", numVertices, numCells);
    return true;
}

void Mesh3D::prepareMesh() {
    i2_context *ctx = i2host::context();
    checkI2Errors(i2_mesh_geometry(ctx, (const double *)vertices.data, vertices.size, (const int *)cells.data, cells.size,
                                   (double *)cellNormals.data, (double *)cellCenters.data, cellMeasures.data));
    long long counts[3] = {0, 0, 0};
    checkI2Errors(i2_classify_count(ctx, (const int *)cells.data, cells.size, counts));
    for (int k = 0; k < 3; ++k) {
        if (counts[k] > 0x3fffffff) checkI2Errors(I2_E_TOOBIG);
        pairCounts[k] = counts[k];
    }
    // multi-GPU runs never need the whole regular list on one device: it is built on first use of getNotNeighbors()
    notNeighborsDeferred = i2host::gpus() > 1;
    for (int k = 0; k < 3; ++k)
        if (counts[k] && !(k == 2 && notNeighborsDeferred)) pairLists[k].allocate((int)counts[k]);
    checkI2Errors(i2_classify_fill(ctx, (const int *)cells.data, cells.size, (int *)pairLists[0].data, (int *)pairLists[1].data,
                                   notNeighborsDeferred ? nullptr : (int *)pairLists[2].data));
    printf("Found %d pairs of simple neighbors and %d pairs of attached neighbors, %d pairs are not neighbors\n", (int)counts[0],
           (int)counts[1], (int)counts[2]);
    checkCudaErrors(cudaDeviceSynchronize());
}

void Mesh3D::materialiseNotNeighbors() const {
    if (!notNeighborsDeferred) return;
    notNeighborsDeferred = false;
    if (!pairCounts[2]) return;
    i2_context *ctx = i2host::context();
    long long counts[3];
    checkI2Errors(i2_classify_count(ctx, (const int *)cells.data, cells.size, counts));
    pairLists[2].allocate((int)counts[2]);
    checkI2Errors(i2_classify_fill(ctx, (const int *)cells.data, cells.size, nullptr, nullptr, (int *)pairLists[2].data));
    checkCudaErrors(cudaDeviceSynchronize());
}

void exportMeshToObj(const std::string &filename, const std::vector<Point3> &vertices, const std::vector<int3> &cells) {
    std::ofstream out(filename.c_str());
    for (const Point3 &p : vertices) out << "v " << p.x << " " << p.y << " " << p.z << std::endl;
    for (const int3 &t : cells) out << "f " << t.x + 1 << " " << t.y + 1 << " " << t.z + 1 << std::endl;
    out.close();
    printf("Mesh saved to %s\n", filename.c_str());
}

// VTK PolyData (.vtp, ascii) with the same element order and attribute names as the reference's writer
// (src/Mesh3d.cu:277-343), so files can be diffed.
void exportMeshToVtk(const std::string &filename, const std::vector<Point3> &vertices, const std::vector<int3> &cells,
                     const std::array<std::vector<unsigned char>, 3> &refinementsRequired) {
    std::ofstream out(filename.c_str());
    out << "<?xml version=\"1.0\" ?> " << std::endl;
    out << "<VTKFile type=\"PolyData\" version=\"0.1\" byte_order=\"LittleEndian\">" << std::endl;
    out << "  <PolyData>" << std::endl;
    out << "    <Piece NumberOfPoints=\"" << vertices.size() << "\" NumberOfPolys=\"" << cells.size() << "\">" << std::endl;

    out << "      <Points>" << std::endl;
    out << "        <DataArray type=\"Float32\" NumberOfComponents=\"3\" Format=\"ascii\">" << std::endl;
    out << "        ";
    for (const Point3 &p : vertices) out << p.x << " " << p.y << " " << p.z << " ";
    out << std::endl;
    out << "        </DataArray>" << std::endl;
    out << "      </Points>" << std::endl;

    out << "      <Polys>" << std::endl;
    out << "        <DataArray type=\"Int32\" Name=\"connectivity\" Format=\"ascii\">" << std::endl;
    out << "          ";
    for (const int3 &t : cells) out << t.x << " " << t.y << " " << t.z << " ";
    out << std::endl;
    out << "        </DataArray>" << std::endl;
    out << "        <DataArray type=\"Int32\" Name=\"offsets\" Format=\"ascii\">" << std::endl;
    out << "          ";
    for (size_t k = 0; k < cells.size(); ++k) out << (k + 1) * 3 << " ";
    out << std::endl;
    out << "        </DataArray>" << std::endl;
    out << "      </Polys>" << std::endl;

    bool any = false;
    for (const auto &r : refinementsRequired) any = any || !r.empty();
    if (any) {
        out << "      <CellData>" << std::endl;
        for (int k = 0; k < 3; ++k) {
            if (refinementsRequired[k].empty()) continue;
            const std::string field = neighborTypeString(neighbour_type_enum(k)) + "Refinements";
            out << "        <DataArray type=\"Int32\" Name=\"" + field + "\" Format=\"ascii\">" << std::endl;
            out << "          ";
            for (size_t c = 0; c < cells.size(); ++c) out << (int)refinementsRequired[k][c] << " ";
            out << std::endl;
            out << "        </DataArray>" << std::endl;
        }
        out << "      </CellData>" << std::endl;
    }
    out << "    </Piece>" << std::endl;
    out << "  </PolyData>" << std::endl;
    out << "</VTKFile>" << std::endl;
    out.close();
    printf("Mesh saved to %s\n", filename.c_str());
}

std::string neighborTypeString(neighbour_type_enum neighborType) {
    static const char *names[3] = {"SimpleNeighbors", "AttachedNeighbors", "NotNeighbors"};
    const int k = (int)neighborType;
    return (k >= 0 && k < 3) ? std::string(names[k]) : std::string();
}
