// integrator2test3D — command-line front end with the reference CLI's flags, stdout lines and export files
// (/root/reference/tests/integrator3D/main.cu:57-185, README flag table).  The reference parses its flags with the
// vendored "Lean Mean C++ Option Parser"; this file carries a small parser for the same ten options
// (short, long, --long=value and "--long value" / "-f value" / "-fvalue" spellings).
#include <array>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/integrator2/Mesh3d.cuh"
#include "../../include/integrator2/NumericalIntegrator3d.cuh"
#include "../../include/integrator2/evaluators/evaluatorJ3DK.cuh"

namespace {

struct Options {
    bool help = false, toObj = false, toVtk = false, toCsv = false, toText = false, toBinary = false, check = false;
    bool haveMesh = false, haveScale = false, haveRefine = false;
    std::string meshfile, scale, refine;
};

struct Spec { char shortName; const char *longName; int kind; };  // kind: 0 flag, 1 non-empty argument, 2 numeric argument
const Spec kSpecs[] = {{'h', "help", 0}, {'f', "meshfile", 1}, {'s', "scale", 2}, {0, "exporttoobj", 0}, {0, "exporttovtk", 0},
                       {0, "exporttocsv", 0}, {0, "exportresults", 0}, {'r', "refine", 2}, {'c', "checkresults", 0},
                       {0, "exportbinary", 0}};   // the last one is not in the reference: full-precision records (Evaluator3D::outputResultsToFile)

void printUsage() {
    std::cout << "USAGE: integrator2test3D [options]\n\nOptions:\n"
                 "   -h,         --help             Print usage and exit.\n"
                 "   -f <arg>,   --meshfile=<arg>   Input mesh file name.\n"
                 "   -s <arg>,   --scale=<arg>      Mesh scale factor.\n"
                 "               --exporttoobj      Export original and refined meshes to OBJ files.\n"
                 "               --exporttovtk      Export original and refined meshes to VTK (VTP) files.\n"
                 "               --exporttocsv      Export results of integration to csv files.\n"
                 "               --exportresults    Export results of integration to text files.\n"
                 "   -r <arg>,   --refine=<arg>     Refine the whole mesh N times.\n"
                 "   -c,         --checkresults     Check correctness of pairs of results.\n"
                 "               --exportbinary     Export results of integration to binary files (full precision; not in the reference).\n"
                 "Environment: I2_GPUS=N shards the run over N GPUs; I2_GATHER=1 gathers the results on device 0; I2_SUMMARY_JSON=file writes a run summary.\n";
}

bool isNumeric(const std::string &s) {
    if (s.empty()) return false;
    char *end = nullptr;
    strtod(s.c_str(), &end);
    return end != s.c_str() && *end == 0;
}

bool store(Options &o, const Spec &sp, const std::string &name, const char *value) {
    if (sp.kind == 0) {
        if (!strcmp(sp.longName, "help")) o.help = true;
        else if (!strcmp(sp.longName, "exporttoobj")) o.toObj = true;
        else if (!strcmp(sp.longName, "exporttovtk")) o.toVtk = true;
        else if (!strcmp(sp.longName, "exporttocsv")) o.toCsv = true;
        else if (!strcmp(sp.longName, "exportresults")) o.toText = true;
        else if (!strcmp(sp.longName, "exportbinary")) o.toBinary = true;
        else o.check = true;
        return true;
    }
    if (sp.kind == 1) {
        if (!value || !*value) { fprintf(stderr, "Option '%s' requires a non-empty argument\n", name.c_str()); return false; }
        o.haveMesh = true; o.meshfile = value;
        return true;
    }
    if (!value || !isNumeric(value)) { fprintf(stderr, "Option '%s' requires a numeric argument\n", name.c_str()); return false; }
    if (!strcmp(sp.longName, "scale")) { o.haveScale = true; o.scale = value; }
    else { o.haveRefine = true; o.refine = value; }
    return true;
}

bool parse(int argc, char **argv, Options &o) {
    for (int a = 0; a < argc; ++a) {
        const std::string arg = argv[a];
        if (arg.size() < 2 || arg[0] != '-') break;  // first non-option ends option parsing
        if (arg == "--") break;
        const Spec *sp = nullptr;
        const char *value = nullptr;
        std::string name = arg;
        if (arg[1] == '-') {
            const size_t eq = arg.find('=');
            const std::string longName = arg.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            name = "--" + longName;
            for (const Spec &s : kSpecs) if (longName == s.longName) sp = &s;
            if (!sp) { fprintf(stderr, "Unknown option '%s'\n", name.c_str()); return false; }
            if (sp->kind) {
                if (eq != std::string::npos) value = argv[a] + eq + 1;
                else if (a + 1 < argc) value = argv[++a];
            }
        } else {
            for (const Spec &s : kSpecs) if (s.shortName && arg[1] == s.shortName) sp = &s;
            name = arg.substr(0, 2);
            if (!sp) { fprintf(stderr, "Unknown option '%s'\n", name.c_str()); return false; }
            if (sp->kind) {
                if (arg.size() > 2) value = argv[a] + 2;
                else if (a + 1 < argc) value = argv[++a];
            }
        }
        if (!store(o, *sp, name, value)) return false;
    }
    return true;
}

}  // namespace

int main(int argc, char *argv[]) {
    argc -= 1;
    argv += 1;
    Options opt;
    if (!parse(argc, argv, opt)) return EXIT_FAILURE;
    if (opt.help || argc == 0) {
        printUsage();
        return EXIT_SUCCESS;
    }
    if (!opt.haveMesh) {
        printf("No input file with mesh specified. Exiting\n");
        return EXIT_FAILURE;
    }
    const double scale = opt.haveScale ? std::stod(opt.scale) : 1.0;

    Mesh3D mesh;
    if (!mesh.loadMeshFromFile(opt.meshfile, scale)) return EXIT_FAILURE;
    mesh.prepareMesh();

    NumericalIntegrator3D numIntegrator(mesh, qf3D13);
    EvaluatorJ3DK evaluator(mesh, numIntegrator);

    int refineLevel = -1;
    if (opt.haveRefine) {
        refineLevel = std::stoi(opt.refine);
        numIntegrator.setFixedRefinementLevel(refineLevel);
        if (refineLevel) printf("Using fixed refinement level equal to %d\n", refineLevel);
        else printf("Using original mesh without refinement\n");
    } else
        printf("Using adaptive error control procedure\n");

    std::vector<Point3> vertices;
    std::vector<int3> cells;
    auto fetch = [&](const deviceVector<Point3> &v, const deviceVector<int3> &c) {
        vertices.resize(v.size);
        cells.resize(c.size);
        copy_d2h(v.data, vertices.data(), v.size);
        copy_d2h(c.data, cells.data(), c.size);
        checkCudaErrors(cudaDeviceSynchronize());
    };

    if (opt.toObj || opt.toVtk) {
        fetch(mesh.getVertices(), mesh.getCells());
        if (opt.toObj) exportMeshToObj("OriginalMesh.obj", vertices, cells);
        if (opt.toVtk && opt.haveRefine) exportMeshToVtk("OriginalMesh.vtp", vertices, cells, {});
    }

    evaluator.runAllPairs(opt.check);

    if (opt.toText || opt.toCsv) {
        const output_format_enum format = opt.toCsv ? output_format_enum::csv : output_format_enum::plainText;  // csv wins
        evaluator.outputResultsToFile(neighbour_type_enum::simple_neighbors, format);
        evaluator.outputResultsToFile(neighbour_type_enum::attached_neighbors, format);
        evaluator.outputResultsToFile(neighbour_type_enum::not_neighbors, format);
    }

    if (opt.toBinary) {
        evaluator.outputResultsToFile(neighbour_type_enum::simple_neighbors, output_format_enum::binary);
        evaluator.outputResultsToFile(neighbour_type_enum::attached_neighbors, output_format_enum::binary);
        evaluator.outputResultsToFile(neighbour_type_enum::not_neighbors, output_format_enum::binary);
    }
    if (const char *path = getenv("I2_SUMMARY_JSON")) {
        if (FILE *f = fopen(path, "w")) {
            fprintf(f, "%s\n", evaluator.getRunSummaryJson().c_str());
            fclose(f);
        }
    }

    if ((opt.toObj || opt.toVtk) && refineLevel > 0) {
        fetch(numIntegrator.getRefinedVertices(), numIntegrator.getRefinedCells());
        if (opt.toObj) exportMeshToObj("RefinedMesh.obj", vertices, cells);
        if (opt.toVtk) exportMeshToVtk("RefinedMesh.vtp", vertices, cells, {});
    }

    if (opt.toVtk && !opt.haveRefine) {
        std::array<std::vector<unsigned char>, 3> refinements;
        for (int k = 0; k < 3; ++k) {
            const auto *counters = numIntegrator.getRefinementsRequired(neighbour_type_enum(k));
            if (counters->size) {
                refinements[k].resize(counters->size);
                copy_d2h(counters->data, refinements[k].data(), counters->size);
            }
        }
        checkCudaErrors(cudaDeviceSynchronize());
        exportMeshToVtk("OriginalMesh.vtp", vertices, cells, refinements);
    }
    return EXIT_SUCCESS;
}
