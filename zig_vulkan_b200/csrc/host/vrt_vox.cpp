// vrt_vox.cpp — MagicaVoxel .vox (v150) reader: C++ counterpart of src/modules/voxel_rt/vox/loader.zig + types.zig,
// plus the two loops of src/main.zig that turn a model into materials and Grid.insert calls (:96-118).
// Unlike the reference ("pos will cause out of bounds easily", loader.zig:88) every read is bounds-checked and a
// truncated file is VRT_VOX_E_INVALID_FILE_CONTENT.
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "vrt_host_internal.h"

struct vrt_vox {
    int32_t num_models = 1;  // Chunk.Pack.num_models
    std::vector<int32_t> sizes;                 // 3 per model (Chunk.Size)
    std::vector<std::vector<uint8_t>> xyzi;     // 4 bytes per voxel (Chunk.XyziElement)
    uint8_t rgba[256 * 4];                      // Chunk.RgbaElement x 256
};

namespace {

bool tag_is(const uint8_t* p, const char* tag) { return std::memcmp(p, tag, 4) == 0; }
int32_t rd_i32(const uint8_t* p) {
    int32_t v;
    std::memcpy(&v, p, 4);  // parseI32 (loader.zig:199-202): little-endian host
    return v;
}

// The format's default palette (loader.zig:246-263 lists it as 256 constants 0xAABBGGRR): index 0 = 0; 1..215 = the
// 6x6x6 colour cube ff,cc,99,66,33,00 with blue fastest and red slowest, minus black; then 10-step ramps
// ee,dd,bb,aa,88,77,55,44,22,11 of red, green, blue and grey.
void default_palette(uint8_t out[256 * 4]) {
    std::memset(out, 0, 256 * 4);
    static const uint8_t lv[6] = {0xff, 0xcc, 0x99, 0x66, 0x33, 0x00};
    for (int i = 0; i < 215; i++) {
        uint8_t* e = out + 4 * (i + 1);
        e[0] = lv[i / 36], e[1] = lv[(i / 6) % 6], e[2] = lv[i % 6], e[3] = 0xff;
    }
    static const uint8_t ramp[10] = {0xee, 0xdd, 0xbb, 0xaa, 0x88, 0x77, 0x55, 0x44, 0x22, 0x11};
    for (int c = 0; c < 4; c++) {
        for (int k = 0; k < 10; k++) {
            uint8_t* e = out + 4 * (216 + c * 10 + k);
            e[3] = 0xff;
            if (c == 3) e[0] = e[1] = e[2] = ramp[k];
            else e[c] = ramp[k];
        }
    }
}

}  // namespace

extern "C" {

// validateHeader (loader.zig:215-228)
int vrt_vox_validate_header(const uint8_t* buffer, size_t len) {
    if (!buffer || len < 12) return VRT_VOX_E_INVALID_FILE_CONTENT;
    if (!tag_is(buffer, "VOX ")) return VRT_VOX_E_INVALID_ID;
    if (buffer[4] != 150) return VRT_VOX_E_UNEXPECTED_VERSION;
    if (!tag_is(buffer + 8, "MAIN")) return VRT_VOX_E_INVALID_FILE_CONTENT;
    return 0;
}

// parseBuffer (loader.zig:41-197)
int vrt_vox_parse(vrt_vox** out, const uint8_t* buffer, size_t len, int strict) {
    if (!out) return VRT_VOX_E_INVALID_FILE_CONTENT;
    *out = nullptr;
    if (!buffer || len < 20) return VRT_VOX_E_INVALID_FILE_CONTENT;
    if (strict) {
        const int rc = vrt_vox_validate_header(buffer, len);
        if (rc) return rc;
    }
    vrt_vox* vox = new (std::nothrow) vrt_vox();
    if (!vox) return VRT_VOX_E_INVALID_FILE_CONTENT;
    const size_t chunk_stride = 12;  // id + chunk size + child size
    size_t pos = 8 + chunk_stride;   // skip the header and the MAIN chunk
    auto need = [&](size_t n) { return pos + n <= len; };
    auto fail = [&](int code) {
        delete vox;
        return code;
    };
    if (!need(1)) return fail(VRT_VOX_E_INVALID_FILE_CONTENT);
    if (buffer[pos] == 'P') {  // optional PACK chunk (:63-76)
        if (!need(chunk_stride + 4)) return fail(VRT_VOX_E_INVALID_FILE_CONTENT);
        pos += chunk_stride;
        vox->num_models = rd_i32(buffer + pos);
        pos += 4;
    }
    if (vox->num_models < 0 || vox->num_models > 1 << 20) return fail(VRT_VOX_E_INVALID_FILE_CONTENT);
    for (int32_t model = 0; model < vox->num_models; model++) {
        if (!need(chunk_stride + 12)) return fail(VRT_VOX_E_INVALID_FILE_CONTENT);
        if (strict && !tag_is(buffer + pos, "SIZE")) return fail(VRT_VOX_E_EXPECTED_SIZE_HEADER);  // :93-97
        pos += chunk_stride;
        for (int i = 0; i < 3; i++) vox->sizes.push_back(rd_i32(buffer + pos + 4 * i));
        pos += 12;
        if (!need(chunk_stride + 4)) return fail(VRT_VOX_E_INVALID_FILE_CONTENT);
        if (strict && !tag_is(buffer + pos, "XYZI")) return fail(VRT_VOX_E_EXPECTED_XYZI_HEADER);  // :120-124
        pos += chunk_stride;
        const int32_t voxel_count = rd_i32(buffer + pos);
        pos += 4;
        if (voxel_count < 0 || !need((size_t)voxel_count * 4)) return fail(VRT_VOX_E_INVALID_FILE_CONTENT);
        vox->xyzi.emplace_back(buffer + pos, buffer + pos + (size_t)voxel_count * 4);
        pos += (size_t)voxel_count * 4;
    }
    bool rgba_set = false;
    while (pos < len) {  // RGBA and extension chunks (:153-190)
        if (buffer[pos] == 'R') {
            if (strict && (!need(4) || !tag_is(buffer + pos, "RGBA"))) return fail(VRT_VOX_E_EXPECTED_RGBA_HEADER);
            if (!need(chunk_stride + 254 * 4)) return fail(VRT_VOX_E_INVALID_FILE_CONTENT);
            pos += chunk_stride;
            vox->rgba[0] = 0, vox->rgba[1] = 0, vox->rgba[2] = 0, vox->rgba[3] = 1;  // :167-172
            std::memcpy(vox->rgba + 4, buffer + pos, 254 * 4);  // palette[1..254]; entry 255 is never written by the reference (:173-182)
            vox->rgba[255 * 4] = vox->rgba[255 * 4 + 1] = vox->rgba[255 * 4 + 2] = vox->rgba[255 * 4 + 3] = 0;
            pos += 254 * 4;
            rgba_set = true;
        } else {
            pos += 4;  // unknown chunk: skipped 4 bytes at a time (:185-188)
        }
    }
    if (!rgba_set) default_palette(vox->rgba);  // :192-195
    *out = vox;
    return 0;
}

// load (loader.zig:9-30); `path` is used as given (the reference joins it to the executable's directory)
int vrt_vox_load(vrt_vox** out, const char* path, int strict) {
    if (!out) return VRT_VOX_E_INVALID_FILE_CONTENT;
    *out = nullptr;
    if (!path) return VRT_VOX_E_IO;
    std::FILE* f = std::fopen(path, "rb");
    if (!f) return VRT_VOX_E_IO;
    std::vector<uint8_t> buf;
    uint8_t chunk[1 << 16];
    size_t n;
    while ((n = std::fread(chunk, 1, sizeof(chunk), f)) > 0) buf.insert(buf.end(), chunk, chunk + n);
    std::fclose(f);
    return vrt_vox_parse(out, buf.data(), buf.size(), strict);
}

void vrt_vox_destroy(vrt_vox* v) { delete v; }
int32_t vrt_vox_num_models(const vrt_vox* v) { return v ? v->num_models : 0; }
int vrt_vox_model_size(const vrt_vox* v, int32_t model, int32_t size_xyz[3]) {
    if (!v || model < 0 || model >= v->num_models || !size_xyz) return -1;
    for (int i = 0; i < 3; i++) size_xyz[i] = v->sizes[(size_t)model * 3 + i];
    return 0;
}
const uint8_t* vrt_vox_model_xyzi(const vrt_vox* v, int32_t model, uint64_t* count) {
    if (!v || model < 0 || model >= v->num_models) return nullptr;
    if (count) *count = v->xyzi[(size_t)model].size() / 4;
    return v->xyzi[(size_t)model].data();
}
const uint8_t* vrt_vox_palette(const vrt_vox* v) { return v ? v->rgba : nullptr; }

// main.zig:96-108: palette entries become materials after the terrain materials; alpha < 0.8 -> dielectric with ir 1.52.
// materials[material_base + i] = rgba[i] for i in [0, capacity - material_base).
uint32_t vrt_vox_materials(const vrt_vox* v, vrt_material* materials, uint32_t capacity, uint32_t material_base) {
    if (!v || !materials || material_base >= capacity) return 0;
    const uint32_t n = capacity - material_base < 256u ? capacity - material_base : 256u;
    for (uint32_t i = 0; i < n; i++) {
        const uint8_t* e = v->rgba + 4 * i;
        vrt_material& m = materials[material_base + i];
        const bool glass = (float)e[3] / 255.0f < 0.8f;
        m.type = glass ? VRT_MAT_DIELECTRIC : VRT_MAT_LAMBERTIAN;
        m.albedo_r = (float)e[0] / 255.0f, m.albedo_g = (float)e[1] / 255.0f, m.albedo_b = (float)e[2] / 255.0f;
        m.type_data = glass ? 1.52f : 0.0f;
    }
    return n;
}

// main.zig:110-118: grid.insert(x + off_x, z + off_y, y + off_z, color_index + material_base) — .vox is z-up.
int vrt_vox_insert_into_grid(const vrt_vox* v, int32_t model, vrt_grid* grid, uint32_t off_x, uint32_t off_y, uint32_t off_z, uint32_t material_base) {
    if (!v || !grid || model < 0 || model >= v->num_models) return -1;
    const std::vector<uint8_t>& e = v->xyzi[(size_t)model];
    for (size_t i = 0; i + 3 < e.size(); i += 4) {
        const int rc = vrt_grid_insert(grid, (uint32_t)e[i] + off_x, (uint32_t)e[i + 2] + off_y, (uint32_t)e[i + 1] + off_z, (uint8_t)(e[i + 3] + material_base));
        if (rc) return rc;
    }
    return 0;
}

}  // extern "C"
