// One-off GPU check of the tile kernels' type table (profiles/r02_summary.md section 5): scores of pre-generated sentences
// over both sides of every CharacterType range edge, through the C ABI, against scores the CPU oracle computed beforehand.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../../include/vaporetto_b200.h"
static std::vector<uint8_t> rd(const char* p) { FILE* f = fopen(p, "rb"); if (!f) { printf("missing %s\n", p); exit(2); } std::vector<uint8_t> b; uint8_t buf[65536]; size_t n; while ((n = fread(buf, 1, sizeof buf, f)) > 0) b.insert(b.end(), buf, buf + n); fclose(f); return b; }
int main(int argc, char** argv) {
    int bad_total = 0;
    for (int m = 1; m < argc; ++m) {
        char path[512];
        snprintf(path, sizeof path, "%s.model", argv[m]); auto model = rd(path);
        snprintf(path, sizeof path, "%s.text", argv[m]); auto text = rd(path);
        snprintf(path, sizeof path, "%s.offs", argv[m]); auto offs_b = rd(path);
        snprintf(path, sizeof path, "%s.scores", argv[m]); auto want_b = rd(path);
        const uint64_t* offs = reinterpret_cast<const uint64_t*>(offs_b.data());
        const size_t n = offs_b.size() / 8 - 1;
        const int32_t* want = reinterpret_cast<const int32_t*>(want_b.data());
        const size_t nb = want_b.size() / 4;
        vpt_model* mod = nullptr; size_t used = 0;
        if (vpt_model_read(model.data(), model.size(), &mod, &used)) { printf("model_read: %s\n", vpt_last_error()); return 2; }
        vpt_predictor* p = nullptr;
        if (vpt_predictor_new(mod, 0, 0, &p)) { printf("predictor_new: %s\n", vpt_last_error()); return 2; }
        text.resize(text.size() + 64, 0);
        std::vector<int32_t> sc(nb + 16), st(n);
        std::vector<uint8_t> bd(nb + 16);
        std::vector<uint64_t> bo(n + 1);
        uint64_t nbo = 0, nco = 0;
        if (vpt_predict_batch(p, text.data(), offs, n, sc.data(), bd.data(), nb + 16, bo.data(), st.data(), nullptr, nullptr, 0, nullptr, &nbo, &nco)) {
            printf("predict_batch: %s\n", vpt_last_error()); return 2; }
        int bad = 0;
        if (nbo != nb) { printf("%s: boundary count %llu, want %zu\n", argv[m], (unsigned long long)nbo, nb); bad = 1; }
        for (size_t i = 0; i < nb && i < nbo; ++i) if (sc[i] != want[i]) { if (bad < 5) printf("%s: score %zu = %d, want %d\n", argv[m], i, sc[i], want[i]); ++bad; }
        for (size_t i = 0; i < n; ++i) if (st[i] != 0) { if (bad < 5) printf("%s: status[%zu] = %d\n", argv[m], i, st[i]); ++bad; }
        printf("EDGE %s: %zu sentences, %zu boundaries, %d mismatches\n", argv[m], n, nb, bad);
        bad_total += bad;
        vpt_predictor_free(p);
    }
    printf(bad_total ? "EDGE FAILED\n" : "EDGE OK\n");
    return bad_total ? 1 : 0;
}
