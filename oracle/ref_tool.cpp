// TEST INFRASTRUCTURE ONLY. Entry point of oracle/_ref/spumoni_ref: the reference's own translation units, compiled
// where they lie under /root/reference (recipe: oracle/Makefile.ref) against the third-party stand-ins of oracle/shim/.
//   spumoni_ref run ...                 the reference's `spumoni run` (spumoni.cpp main, renamed at compile time)
//   spumoni_ref x-build-pml <ref.fa>    build_spumoni_main     (compute_ms_pml.cpp:1397): .bwt.heads/.bwt.len/.thr_pos -> .thrbv.spumoni
//   spumoni_ref x-build-ms  <ref.fa>    build_spumoni_ms_main  (compute_ms_pml.cpp:1384): + .ssa/.esa -> .thrbv.ms
//   spumoni_ref x-doc <ref.fa> <runs>   DocumentArray(ref, runs) + serialize (doc_array.cpp:18-108, :184-195) -> .doc
//   spumoni_ref x-nulldb <ref.fa> <null_reads.fa> pml|ms <bin>   the null-database step of `spumoni build` (spumoni.cpp:654-702)
// The x- commands expose the build-side functions whose helper programs (PFP, thresholds, RePair) are not available.
#include <compute_ms_pml.hpp>
#include <doc_array.hpp>
#include <emp_null_database.hpp>
#include <spumoni_main.hpp>

#include <cstring>
#include <fstream>
#include <string>

int spumoni_reference_main(int argc, char** argv);

int main(int argc, char** argv) {
    const std::string cmd = argc > 1 ? argv[1] : "";
    if (cmd == "x-build-pml" && argc == 3) {
        auto st = build_spumoni_main(argv[2], argv[2]);
        std::printf("n %zu r %zu\n", st.first, st.second);
        return 0;
    }
    if (cmd == "x-build-ms" && argc == 3) {
        auto st = build_spumoni_ms_main(argv[2], argv[2]);
        std::printf("n %zu r %zu\n", st.first, st.second);
        return 0;
    }
    if (cmd == "x-doc" && argc == 4) {
        DocumentArray doc(argv[2], std::stoull(argv[3]));
        std::ofstream out(std::string(argv[2]) + ".doc");
        doc.serialize(out);
        return 0;
    }
    if (cmd == "x-nulldb" && argc == 6) {
        const bool ms = std::strcmp(argv[4], "ms") == 0;
        EmpNullDatabase db(argv[2], argv[3], false, ms ? MS : PML, false, false, 4, 11, false);
        if (ms) find_threshold_based_on_null_ms_distribution(argv[2], argv[3], false, false, false, 4, 11, db, std::stoull(argv[5]));
        else find_threshold_based_on_null_pml_distribution(argv[2], argv[3], false, false, false, 4, 11, db, std::stoull(argv[5]));
        std::ofstream out(std::string(argv[2]) + (ms ? ".msnulldb" : ".pmlnulldb"));
        db.serialize(out);
        std::printf("num_values %zu ks_stat_threshold %.17g mean %.17g percentile %.17g\n", db.num_values, db.ks_stat_threshold,
                    db.mean_null_stat, db.percentile_value);
        return 0;
    }
    return spumoni_reference_main(argc, argv);
}
