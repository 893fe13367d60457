// mmpld.cpp -- MMPLD frame ingest into pinned, double-buffered host memory.
//
// Replaces, for the B200 path, the frame loader of moldyn::MMPLDDataSource: header + seek table
// (plugins/moldyn/src/io/MMPLDDataSource.cpp:375-401) and the per-frame list parser (Frame::SetData, :61-217; format
// also in utils/MMPLD/mmpldinfo.py:69-120).  The reference hands out pointers into an unpinned vislib::RawStorage
// filled by a loader thread; here a frame is read straight into cudaHostAlloc'ed memory (a plain DMA source for
// mms_push_particles) and described as an array of mms_list -- the same (pointer, type, stride) view a
// MultiParticleDataCall carries.  A background thread can prefetch the next frame into the second buffer.
// Host-only code: no kernels.  If pinning fails (no CUDA device) the buffers are ordinary aligned memory.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/mmsurf.h"

namespace {

const unsigned kVrtSize[5] = {0, 12, 16, 6, 24};
// FILE colour type -> (bytes, in-memory ColourDataType); the file enum differs from the in-memory one (:118-151)
const unsigned kColSize[8] = {0, 3, 4, 4, 12, 16, 8, 8};
const int kColMem[8] = {MMS_COL_NONE, MMS_COL_UINT8_RGB, MMS_COL_UINT8_RGBA, MMS_COL_FLOAT_I, MMS_COL_FLOAT_RGB,
    MMS_COL_FLOAT_RGBA, MMS_COL_USHORT_RGBA, MMS_COL_DOUBLE_I};

struct Buffer {
    char* base = nullptr;
    size_t cap = 0;
    bool pinned = false;
    uint32_t frame = UINT32_MAX;
    size_t pad = 0; // frame bytes start at base + pad (chosen so that the first payload is 16-byte aligned)
    std::vector<mms_list> lists;
    float timestamp = 0.0f;
    bool ok = false;
    std::string err;
    void release() {
        if (base) {
            if (pinned) cudaFreeHost(base);
            else free(base);
        }
        base = nullptr;
        cap = 0;
    }
    bool ensure(size_t bytes) {
        if (bytes <= cap) return true;
        release();
        const size_t want = bytes + bytes / 4;
        void* p = nullptr;
        if (cudaHostAlloc(&p, want, cudaHostAllocDefault) == cudaSuccess) {
            pinned = true;
        } else {
            cudaGetLastError();
            pinned = false;
            if (posix_memalign(&p, 256, want) != 0) p = nullptr;
        }
        base = static_cast<char*>(p);
        cap = base ? want : 0;
        return base != nullptr;
    }
};

} // namespace

struct mms_mmpld {
    FILE* f = nullptr;
    std::string path, err;
    uint16_t version = 0;
    uint32_t frames = 0;
    float bbox[6]{}, clip[6]{};
    std::vector<uint64_t> seek;
    Buffer buf[2];
    int cur = 0;
    std::thread loader;
    bool loading = false;
    uint32_t loadingFrame = 0;
};

namespace {

template<class T> T rd(const char* p) {
    T v;
    std::memcpy(&v, p, sizeof(T));
    return v;
}

/** Reads frame `frame` into `b` and parses its list table.  Uses its own FILE handle position via pread-like seek. */
bool loadFrame(mms_mmpld* m, uint32_t frame, Buffer& b) {
    b.ok = false;
    b.lists.clear();
    if (frame >= m->frames) {
        b.err = "frame index out of range";
        return false;
    }
    const uint64_t beg = m->seek[frame], end = m->seek[frame + 1];
    if (end < beg) {
        b.err = "corrupt seek table";
        return false;
    }
    const size_t bytes = static_cast<size_t>(end - beg);
    if (!b.ensure(bytes + 64)) {
        b.err = "out of host memory";
        return false;
    }
    // the first list's payload starts after: [timestamp] + list count + list header; compute it from a peek of the header
    FILE* f = fopen(m->path.c_str(), "rb");
    if (!f) {
        b.err = "cannot reopen file";
        return false;
    }
    char head[64];
    const size_t headBytes = bytes < sizeof(head) ? bytes : sizeof(head);
    bool ok = fseeko(f, static_cast<off_t>(beg), SEEK_SET) == 0 && fread(head, 1, headBytes, f) == headBytes;
    size_t firstPayload = 0;
    if (ok && bytes >= 8) {
        size_t p = m->version >= 102 ? 4 : 0;
        p += 4; // list count
        if (p + 2 <= headBytes) {
            const uint8_t vt = static_cast<uint8_t>(head[p]), ct = static_cast<uint8_t>(head[p + 1]);
            p += 2;
            if (vt == 1 || vt == 3 || vt == 4) p += 4;
            if (ct == 0) p += 4;
            else if (ct == 3 || ct == 7) p += 8;
            p += 8;
            if (m->version >= 103) p += 24;
            firstPayload = p;
        }
    }
    b.pad = (16 - (firstPayload & 15)) & 15;
    char* dat = b.base + b.pad;
    ok = ok && fseeko(f, static_cast<off_t>(beg), SEEK_SET) == 0 && fread(dat, 1, bytes, f) == bytes;
    fclose(f);
    if (!ok) {
        b.err = "short read";
        return false;
    }
    size_t p = 0;
    auto need = [&](size_t n) { return p + n <= bytes; };
    b.timestamp = static_cast<float>(frame);
    if (m->version >= 102) {
        if (!need(4)) { b.err = "truncated frame"; return false; }
        b.timestamp = rd<float>(dat + p);
        p += 4;
    }
    if (!need(4)) { b.err = "truncated frame"; return false; }
    const uint32_t plc = rd<uint32_t>(dat + p);
    p += 4;
    for (uint32_t i = 0; i < plc; ++i) {
        if (!need(2)) { b.err = "truncated list header"; return false; }
        const uint8_t vt = static_cast<uint8_t>(dat[p]), ct = static_cast<uint8_t>(dat[p + 1]);
        p += 2;
        mms_list l{};
        const unsigned vsz = vt <= 4 ? kVrtSize[vt] : 0;
        const unsigned csz = (vt != 0 && vt <= 4 && ct <= 7) ? kColSize[ct] : 0;
        l.vtx_type = vt <= 4 ? vt : 0;
        l.col_type = csz ? kColMem[ct] : MMS_COL_NONE;
        l.global_radius = 0.05f;
        if (vt == 1 || vt == 3 || vt == 4) {
            if (!need(4)) { b.err = "truncated list header"; return false; }
            l.global_radius = rd<float>(dat + p);
            p += 4;
        }
        l.global_rgba[0] = l.global_rgba[1] = l.global_rgba[2] = 192, l.global_rgba[3] = 255;
        l.irange[0] = 0.0f, l.irange[1] = 1.0f;
        if (ct == 0) {
            if (!need(4)) { b.err = "truncated list header"; return false; }
            for (int k = 0; k < 3; ++k) l.global_rgba[k] = static_cast<uint8_t>(dat[p + k]); // alpha stays 255 like SetGlobalColour(r,g,b)
            p += 4;
        } else if (ct == 3 || ct == 7) {
            if (!need(8)) { b.err = "truncated list header"; return false; }
            l.irange[0] = rd<float>(dat + p), l.irange[1] = rd<float>(dat + p + 4);
            p += 8;
        }
        if (!need(8)) { b.err = "truncated list header"; return false; }
        l.count = rd<uint64_t>(dat + p);
        p += 8;
        if (m->version >= 103) {
            if (!need(24)) { b.err = "truncated list header"; return false; }
            p += 24; // per-list bounding box (not needed by the density path)
        }
        const unsigned stride = vsz + csz;
        l.vtx_stride = l.col_stride = stride;
        l.vtx = dat + p;
        l.col = csz ? dat + p + vsz : nullptr;
        const size_t payload = static_cast<size_t>(stride) * l.count;
        if (!need(payload)) { b.err = "truncated particle payload"; return false; }
        p += payload;
        if (m->version == 101) { // cluster infos trailer (:204-215)
            if (!need(4 + sizeof(size_t))) { b.err = "truncated cluster info"; return false; }
            p += 4;
            const size_t sz = rd<size_t>(dat + p);
            p += sizeof(size_t);
            if (!need(sz)) { b.err = "truncated cluster info"; return false; }
            p += sz;
        }
        b.lists.push_back(l);
    }
    b.frame = frame;
    b.ok = true;
    return true;
}

void joinLoader(mms_mmpld* m) {
    if (m->loader.joinable()) m->loader.join();
    m->loading = false;
}

} // namespace

extern "C" {

int mms_mmpld_open(mms_mmpld** out, const char* path) {
    if (!out || !path) return MMS_ERR_INVALID;
    *out = nullptr;
    auto* m = new mms_mmpld();
    m->path = path;
    FILE* f = fopen(path, "rb");
    auto fail = [&](const char* msg) {
        if (f) fclose(f);
        // keep the object so that the caller can read the message
        m->err = msg;
        *out = m;
        return MMS_ERR_INVALID;
    };
    if (!f) return fail("unable to open MMPLD file");
    char magic[6];
    if (fread(magic, 1, 6, f) != 6 || std::memcmp(magic, "MMPLD\0", 6) != 0) return fail("MMPLD file header id wrong");
    if (fread(&m->version, 2, 1, f) != 1 || m->version < 100 || m->version > 103) return fail("MMPLD file header version wrong");
    if (fread(&m->frames, 4, 1, f) != 1 || m->frames == 0) return fail("MMPLD file does not contain any frame information");
    if (fread(m->bbox, 4, 6, f) != 6 || fread(m->clip, 4, 6, f) != 6) return fail("unable to read MMPLD file header");
    m->seek.resize(static_cast<size_t>(m->frames) + 1);
    if (fread(m->seek.data(), 8, m->seek.size(), f) != m->seek.size()) return fail("unable to read MMPLD seek table");
    fclose(f);
    *out = m;
    return MMS_OK;
}

int mms_mmpld_close(mms_mmpld* m) {
    if (!m) return MMS_ERR_INVALID;
    joinLoader(m);
    m->buf[0].release();
    m->buf[1].release();
    delete m;
    return MMS_OK;
}

const char* mms_mmpld_last_error(const mms_mmpld* m) { return m ? m->err.c_str() : "null reader"; }

int mms_mmpld_info(const mms_mmpld* m, uint32_t* frames, uint32_t* version, float bbox[6], float clipbox[6]) {
    if (!m || m->frames == 0) return MMS_ERR_INVALID;
    if (frames) *frames = m->frames;
    if (version) *version = m->version;
    if (bbox) std::memcpy(bbox, m->bbox, sizeof(m->bbox));
    if (clipbox) std::memcpy(clipbox, m->clip, sizeof(m->clip));
    return MMS_OK;
}

int mms_mmpld_prefetch(mms_mmpld* m, uint32_t frame) {
    if (!m || m->frames == 0) return MMS_ERR_INVALID;
    joinLoader(m);
    Buffer& b = m->buf[m->cur ^ 1];
    if (b.ok && b.frame == frame) return MMS_OK;
    m->loading = true;
    m->loadingFrame = frame;
    m->loader = std::thread([m, frame, &b]() { loadFrame(m, frame, b); });
    return MMS_OK;
}

int mms_mmpld_read_frame(mms_mmpld* m, uint32_t frame, int32_t* nlists, const mms_list** lists, float* timestamp) {
    if (!m || m->frames == 0 || !nlists || !lists) return MMS_ERR_INVALID;
    joinLoader(m);
    Buffer* b = &m->buf[m->cur];
    if (!(b->ok && b->frame == frame)) {
        Buffer& other = m->buf[m->cur ^ 1];
        if (other.ok && other.frame == frame) { // prefetched
            m->cur ^= 1;
            b = &other;
        } else {
            m->cur ^= 1; // keep the previously returned frame valid for one more call
            b = &m->buf[m->cur];
            if (!loadFrame(m, frame, *b)) {
                m->err = b->err;
                return MMS_ERR_INVALID;
            }
        }
    }
    *nlists = static_cast<int32_t>(b->lists.size());
    *lists = b->lists.data();
    if (timestamp) *timestamp = b->timestamp;
    return MMS_OK;
}

} // extern "C"
