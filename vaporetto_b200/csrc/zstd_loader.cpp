// vaporetto_b200 — zstd-compressed model images (the reference CLI reads its model through zstd::Decoder:
// predict/src/main.rs:110-111; the released *.model.zst files).  libzstd is not linked: libzstd.so.1 is opened at run
// time (dlopen) and only its streaming decoder is used, so that frames without a stored content size and concatenated
// frames decode the way zstd::Decoder reads them.
#include "zstd_loader.hpp"

#include <dlfcn.h>

#include <mutex>

namespace vpt {

namespace {

struct ZBufIn { const void* src; size_t size; size_t pos; };
struct ZBufOut { void* dst; size_t size; size_t pos; };

struct ZstdApi {
    void* handle = nullptr;
    void* (*create)() = nullptr;
    size_t (*free_)(void*) = nullptr;
    size_t (*init)(void*) = nullptr;
    size_t (*run)(void*, ZBufOut*, ZBufIn*) = nullptr;
    unsigned (*is_error)(size_t) = nullptr;
    const char* (*error_name)(size_t) = nullptr;
    std::string why;
};

const ZstdApi& zstd_api() {
    static ZstdApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libzstd.so.1", "libzstd.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.why = "libzstd.so.1 cannot be opened"; return; }
        auto sym = [&](const char* n) { return dlsym(api.handle, n); };
        api.create = reinterpret_cast<void* (*)()>(sym("ZSTD_createDStream"));
        api.free_ = reinterpret_cast<size_t (*)(void*)>(sym("ZSTD_freeDStream"));
        api.init = reinterpret_cast<size_t (*)(void*)>(sym("ZSTD_initDStream"));
        api.run = reinterpret_cast<size_t (*)(void*, ZBufOut*, ZBufIn*)>(sym("ZSTD_decompressStream"));
        api.is_error = reinterpret_cast<unsigned (*)(size_t)>(sym("ZSTD_isError"));
        api.error_name = reinterpret_cast<const char* (*)(size_t)>(sym("ZSTD_getErrorName"));
        if (!api.create || !api.free_ || !api.init || !api.run || !api.is_error || !api.error_name) {
            api.why = "libzstd.so.1 lacks the streaming decoder";
            api.handle = nullptr;
        }
    });
    return api;
}

}  // namespace

bool is_zstd_frame(const uint8_t* data, size_t len) {
    if (len < 4) return false;
    const uint32_t magic = uint32_t(data[0]) | (uint32_t(data[1]) << 8) | (uint32_t(data[2]) << 16) | (uint32_t(data[3]) << 24);
    return magic == 0xFD2FB528u || (magic & 0xFFFFFFF0u) == 0x184D2A50u;  // a frame or a skippable frame
}

std::vector<uint8_t> zstd_decode_all(const uint8_t* data, size_t len) {
    const ZstdApi& z = zstd_api();
    if (!z.handle) throw Error(kIoError, "IOError: zstd: " + z.why);
    void* ds = z.create();
    if (!ds) throw Error(kIoError, "IOError: zstd: cannot create a decoder");
    struct Guard { const ZstdApi& z; void* p; ~Guard() { z.free_(p); } } guard{z, ds};
    size_t rc = z.init(ds);
    if (z.is_error(rc)) throw Error(kIoError, std::string("IOError: zstd: ") + z.error_name(rc));
    std::vector<uint8_t> out;
    out.resize(len < (1u << 20) ? (1u << 22) : len * 4);
    ZBufIn in{data, len, 0};
    size_t produced = 0;
    size_t last = 1;  // hint of the last call: 0 = a frame ended exactly here
    while (in.pos < in.size) {
        if (produced == out.size()) out.resize(out.size() * 2);
        ZBufOut o{out.data(), out.size(), produced};
        last = z.run(ds, &o, &in);
        if (z.is_error(last)) throw Error(kIoError, std::string("IOError: zstd: ") + z.error_name(last));
        const bool progressed = o.pos != produced;
        produced = o.pos;
        if (!progressed && in.pos == in.size) break;
    }
    // the input is consumed; flush what the decoder still holds
    for (;;) {
        if (produced == out.size()) out.resize(out.size() * 2);
        ZBufOut o{out.data(), out.size(), produced};
        ZBufIn none{data + len, 0, 0};
        const size_t before = produced;
        const size_t r = z.run(ds, &o, &none);
        if (z.is_error(r)) throw Error(kIoError, std::string("IOError: zstd: ") + z.error_name(r));
        produced = o.pos;
        if (produced == before) { if (last != 0 && r != 0) throw Error(kIoError, "IOError: zstd: unexpected end of the frame"); break; }
        last = r;
    }
    out.resize(produced);
    return out;
}

}  // namespace vpt
