// TEST INFRASTRUCTURE — not part of the product.
//
// ref_dump: runs the UNMODIFIED reference library (compiled from /root/reference by
// oracle/build_ref.sh) through its own public API, exactly like the reference CLI does
// (tests/integrator3D/main.cu:102-141), and dumps the per-task results in binary,
// full precision, because the reference's own csv export keeps only 6 significant
// digits (src/evaluators/evaluator3d.cu:415-445).
//
// The only thing this file adds to the reference is a subclass that reads the protected
// device buffers of Evaluator3D (src/evaluators/evaluator3d.cuh:207-236).
//
// usage: ref_dump -f mesh.dat [-s scale] [-r N] -o out_prefix [--pairs file.bin]
//   writes  <out_prefix>.{simple,attached,not}.bin  and <out_prefix>.meta.txt
//   --not-stride K: keep only every K-th record of the (huge) not-neighbours class in the dump
//   --small-stride K: the same for the two adjacent classes
//   --pairs: instead of runAllPairs, run runPairs on 3 user lists read from a binary file
//            (int32 n_simple, n_attached, n_not, then the int3 triples).
//
// record layout of a class file (little endian):
//   int32 n; int32 tasks[n][3]; double results[n][3]; double integrals[n][4];
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "Mesh3d.cuh"
#include "NumericalIntegrator3d.cuh"
#include "evaluators/evaluatorJ3DK.cuh"
#include "common/cuda_memory.cuh"

class DumpEvaluator : public EvaluatorJ3DK
{
public:
    DumpEvaluator(const Mesh3D &mesh_, NumericalIntegrator3D &ni_) : EvaluatorJ3DK(mesh_, ni_) {}

    bool dumpClass(neighbour_type_enum type, const std::string &filename, int stride = 1) const {
        const deviceVector<int3> *tasks = getTasks(type);
        const deviceVector<Point3> *results = nullptr;
        const deviceVector<double4> *integrals = nullptr;
        switch (type) {
        case neighbour_type_enum::simple_neighbors:   results = &d_simpleNeighborsResults;   integrals = &d_simpleNeighborsIntegrals;   break;
        case neighbour_type_enum::attached_neighbors: results = &d_attachedNeighborsResults; integrals = &d_attachedNeighborsIntegrals; break;
        case neighbour_type_enum::not_neighbors:      results = &d_notNeighborsResults;      integrals = &d_notNeighborsIntegrals;      break;
        default: return false;
        }
        int n = tasks->size;
        std::vector<int3> hTasks(n);
        std::vector<Point3> hResults(n);
        std::vector<double4> hIntegrals(n);
        if (n) {
            checkCudaErrors(cudaMemcpy(hTasks.data(), tasks->data, n * sizeof(int3), cudaMemcpyDeviceToHost));
            checkCudaErrors(cudaMemcpy(hResults.data(), results->data, n * sizeof(Point3), cudaMemcpyDeviceToHost));
            checkCudaErrors(cudaMemcpy(hIntegrals.data(), integrals->data, n * sizeof(double4), cudaMemcpyDeviceToHost));
        }
        if (stride > 1) {  // keep records 0, stride, 2*stride, ...
            int m = 0;
            for (int t = 0; t < n; t += stride, ++m) { hTasks[m] = hTasks[t]; hResults[m] = hResults[t]; hIntegrals[m] = hIntegrals[t]; }
            const int full = n;
            n = m;
            (void)full;
        }
        FILE *f = fopen(filename.c_str(), "wb");
        if (!f) return false;
        fwrite(&n, sizeof(int), 1, f);
        fwrite(hTasks.data(), sizeof(int3), n, f);
        fwrite(hResults.data(), sizeof(Point3), n, f);
        fwrite(hIntegrals.data(), sizeof(double4), n, f);
        fclose(f);
        return true;
    }
};

int main(int argc, char **argv)
{
    std::string meshfile, out = "ref", pairsfile;
    double scale = 1.0;
    int refine = -1, notStride = 1, smallStride = 1;
    for (int a = 1; a < argc; ++a) {
        if (!strcmp(argv[a], "-f") && a + 1 < argc) meshfile = argv[++a];
        else if (!strcmp(argv[a], "-s") && a + 1 < argc) scale = atof(argv[++a]);
        else if (!strcmp(argv[a], "-r") && a + 1 < argc) refine = atoi(argv[++a]);
        else if (!strcmp(argv[a], "-o") && a + 1 < argc) out = argv[++a];
        else if (!strcmp(argv[a], "--not-stride") && a + 1 < argc) notStride = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--small-stride") && a + 1 < argc) smallStride = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--pairs") && a + 1 < argc) pairsfile = argv[++a];
        else { fprintf(stderr, "unknown argument %s\n", argv[a]); return 2; }
    }
    if (meshfile.empty()) { fprintf(stderr, "no mesh file\n"); return 2; }

    Mesh3D mesh;
    if (!mesh.loadMeshFromFile(meshfile, scale)) return 1;
    mesh.prepareMesh();

    NumericalIntegrator3D numIntegrator(mesh, qf3D13);
    DumpEvaluator evaluator(mesh, numIntegrator);
    if (refine >= 0) {
        numIntegrator.setFixedRefinementLevel(refine);
        printf("Using fixed refinement level equal to %d\n", refine);
    } else
        printf("Using adaptive error control procedure\n");

    if (pairsfile.empty())
        evaluator.runAllPairs(true);
    else {
        FILE *pf = fopen(pairsfile.c_str(), "rb");
        if (!pf) { fprintf(stderr, "cannot open %s\n", pairsfile.c_str()); return 1; }
        int n[3];
        if (fread(n, sizeof(int), 3, pf) != 3) return 1;
        std::vector<int3> lists[3];
        for (int c = 0; c < 3; ++c) {
            lists[c].resize(n[c]);
            if (n[c] && fread(lists[c].data(), sizeof(int3), n[c], pf) != (size_t)n[c]) return 1;
        }
        fclose(pf);
        evaluator.runPairs(lists[0], lists[1], lists[2]);
    }
    checkCudaErrors(cudaDeviceSynchronize());

    evaluator.dumpClass(neighbour_type_enum::simple_neighbors, out + ".simple.bin", smallStride);
    evaluator.dumpClass(neighbour_type_enum::attached_neighbors, out + ".attached.bin", smallStride);
    evaluator.dumpClass(neighbour_type_enum::not_neighbors, out + ".not.bin", notStride);

    // mesh as the reference sees it + adaptive refinement counters
    {
        const int nv = mesh.getVertices().size, nc = mesh.getCells().size;
        std::vector<Point3> v(nv);
        std::vector<int3> c(nc);
        std::vector<Point3> nrm(nc);
        std::vector<double> area(nc);
        checkCudaErrors(cudaMemcpy(v.data(), mesh.getVertices().data, nv * sizeof(Point3), cudaMemcpyDeviceToHost));
        checkCudaErrors(cudaMemcpy(c.data(), mesh.getCells().data, nc * sizeof(int3), cudaMemcpyDeviceToHost));
        checkCudaErrors(cudaMemcpy(nrm.data(), mesh.getCellNormals().data, nc * sizeof(Point3), cudaMemcpyDeviceToHost));
        checkCudaErrors(cudaMemcpy(area.data(), mesh.getCellMeasures().data, nc * sizeof(double), cudaMemcpyDeviceToHost));
        FILE *f = fopen((out + ".mesh.bin").c_str(), "wb");
        fwrite(&nv, sizeof(int), 1, f);
        fwrite(&nc, sizeof(int), 1, f);
        fwrite(v.data(), sizeof(Point3), nv, f);
        fwrite(c.data(), sizeof(int3), nc, f);
        fwrite(nrm.data(), sizeof(Point3), nc, f);
        fwrite(area.data(), sizeof(double), nc, f);
        int adaptive = refine < 0;
        fwrite(&adaptive, sizeof(int), 1, f);
        if (adaptive)
            for (int k = 0; k < 3; ++k) {
                const auto *rr = numIntegrator.getRefinementsRequired(neighbour_type_enum(k));
                std::vector<unsigned char> h(nc, 0);
                if (rr->size)
                    checkCudaErrors(cudaMemcpy(h.data(), rr->data, nc, cudaMemcpyDeviceToHost));
                fwrite(h.data(), 1, nc, f);
            }
        fclose(f);
    }
    return 0;
}
