// multi_frame_sr_b200 — C++ host of the B200 burst super-resolution path, over the C ABI of include/mfsr.h only
// (no CUDA headers, no torch): what a maintainer of the reference gets by swapping the cv::superres block of
// finalProject/Project/multi_frame_sr.cpp:165-203 for libmfsr_b200.so.
//
//   ./multi_frame_sr_b200 optFlowName inputName iterations          (multi_frame_sr.cpp:122-144: same three arguments)
//
// * inputName city / car / iso select the reference's frame sets (:151-163) with Netpbm files instead of PNG/JPEG — this image
//   has no OpenCV C++ to decode those: img_%06d.{pgm,ppm} (5), car/%d.{pgm,ppm} (4), iso/%06d.{pgm,ppm} (4), numbered from 1
//   like the reference (from 0 is tried too: the repository ships img_000000..4).  Any other inputName is a printf pattern
//   and needs a 4th argument: the number of frames.
//   P5 (maxval > 255: 16-bit big-endian) = Bayer RGGB or gray raw frames, used as they are; P6 8-bit colour frames are
//   mosaiced to RGGB on the 10-bit range of the default parameters (raw = round(v * 959 / 255) + 64).
// * optFlowName is accepted and echoed into the output names (:207-209); the burst path has its own aligner, `iterations`
//   sets its Lucas-Kanade sweeps (:181 setIterations).
// * like the reference, the burst is processed num_times = 10 times and the last real_times = 5 are timed (:146-149,:188-203);
//   prints "<t> sec" and "<fps> FPS" (:204-205), writes <input>_<flow>_sr_result.ppm and the sharpened _sr2_result.ppm (:206-209).
// * `--radius R` (anywhere on the command line) selects the reference's temporal-area pull mode (:182 setTemporalAreaRadius(1),
//   :185-194): the source is the frame set repeated num_times (:168-176), every nextFrame() returns the super-resolved frame i
//   merged from frames [i - R, i + R] (clipped at the ends of the sequence) until the sequence is exhausted, the frames from
//   start_i = (num_times - real_times) * num_images on are timed (:167,:188-190).  Without it the whole frame set is one burst
//   merged onto its first frame (the BASELINE configurations).
#include "mfsr.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct RawFrame {
    int w = 0, h = 0;
    bool gray = false;                 // P5 8-bit input: monochrome, kept as gray
    std::vector<uint16_t> px;          // dense rows
};

bool next_token(FILE* f, std::string& tok)
{
    tok.clear();
    int c;
    while ((c = fgetc(f)) != EOF) {
        if (c == '#') { while ((c = fgetc(f)) != EOF && c != '\n') {} continue; }
        if (c == ' ' || c == '\t' || c == '\n' || c == '\r') { if (!tok.empty()) return true; continue; }
        tok.push_back((char)c);
    }
    return !tok.empty();
}

// Netpbm P5 / P6 -> 16-bit raw frame (see the header comment for the mapping)
bool load_frame(const std::string& path, RawFrame& out)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    std::string magic, sw, sh, smax;
    bool ok = next_token(f, magic) && next_token(f, sw) && next_token(f, sh) && next_token(f, smax);
    const int w = ok ? atoi(sw.c_str()) : 0, h = ok ? atoi(sh.c_str()) : 0, maxv = ok ? atoi(smax.c_str()) : 0;
    ok = ok && w > 0 && h > 0 && maxv > 0 && maxv < 65536 && (magic == "P5" || magic == "P6");
    if (!ok) { fclose(f); return false; }
    const int ch = magic == "P6" ? 3 : 1, bps = maxv > 255 ? 2 : 1;
    std::vector<unsigned char> buf((size_t)w * h * ch * bps);
    ok = fread(buf.data(), 1, buf.size(), f) == buf.size();
    fclose(f);
    if (!ok) return false;
    const int ew = w & ~1, eh = h & ~1;          // the path wants even dimensions
    out.w = ew; out.h = eh; out.gray = false;
    out.px.assign((size_t)ew * eh, 0);
    auto sample = [&](int x, int y, int c) -> int {
        const size_t i = ((size_t)y * w + x) * ch + c;
        return bps == 2 ? (buf[2 * i] << 8) | buf[2 * i + 1] : buf[i];
    };
    for (int y = 0; y < eh; y++)
        for (int x = 0; x < ew; x++) {
            int v;
            if (ch == 1 && bps == 2) v = sample(x, y, 0);                                    // raw frame as it is
            else {
                const int c = ch == 1 ? 0 : ((y & 1) ? ((x & 1) ? 2 : 1) : ((x & 1) ? 1 : 0));   // RGGB: R G / G B
                v = (int)std::floor(sample(x, y, c) * 959.0 / maxv + 0.5) + 64;
            }
            out.px[(size_t)y * ew + x] = (uint16_t)v;
        }
    out.gray = (ch == 1 && bps == 1);
    return true;
}

bool write_ppm(const std::string& path, const std::vector<unsigned char>& rgb, int w, int h)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    fprintf(f, "P6\n%d %d\n255\n", w, h);
    const bool ok = fwrite(rgb.data(), 1, rgb.size(), f) == rgb.size();
    fclose(f);
    return ok;
}

// sharpenImg2 (multi_frame_sr.cpp:90-119): 5 c - left - right - up - down, saturated; the reference advances its output pointer
// from the START of the row while reading from column 1, so the result sits one pixel to the left; border rows / columns are 0
// (the one column the reference leaves uninitialised is 0 here).
std::vector<unsigned char> sharpen(const std::vector<unsigned char>& img, int w, int h)
{
    std::vector<unsigned char> out(img.size(), 0);
    const int ch = 3, row = w * ch;
    for (int y = 1; y < h - 1; y++) {
        const unsigned char* cur = &img[(size_t)y * row];
        unsigned char* o = &out[(size_t)y * row];
        for (int col = ch; col < (w - 1) * ch; col++) {
            const int v = 5 * cur[col] - cur[col - ch] - cur[col + ch] - cur[col - row] - cur[col + row];
            *o++ = (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
        for (int c = 0; c < ch; c++) { out[(size_t)y * row + c] = 0; out[(size_t)y * row + (w - 1) * ch + c] = 0; }
    }
    return out;
}

// The pull-style source of the reference (MultiFrameSource_CUDA, multi_frame_sr.cpp:18-49): hands out one frame per call,
// nullptr when exhausted, reset() rewinds.
class BurstFrameSource {
public:
    // `repeat`: the reference fills its vector with the frame set num_times over (:168-176); here the set is stored once
    explicit BurstFrameSource(std::vector<RawFrame> frames, size_t repeat = 1) : frames_(std::move(frames)), repeat_(repeat) {}
    const RawFrame* nextFrame() { return index_ < size() ? &frames_[index_++ % frames_.size()] : nullptr; }
    void reset() { index_ = 0; }
    size_t size() const { return frames_.size() * repeat_; }
private:
    size_t index_ = 0;
    std::vector<RawFrame> frames_;
    size_t repeat_;
};

// The slice of cv::superres::SuperResolution the reference program uses (:179-194), on top of one mfsr handle.
class BurstSuperResolution {
public:
    ~BurstSuperResolution() { if (h_) mfsr_destroy(h_); }
    void setScale(int s) { scale_ = s; }
    void setIterations(int it) { iterations_ = it; }
    void setTemporalAreaRadius(int r) { radius_ = r; seq_.clear(); pos_ = 0; }      // r < 0: the whole source is one burst
    void setInput(BurstFrameSource* src) { src_ = src; seq_.clear(); pos_ = 0; }
    int outWidth() const { return ow_; }
    int outHeight() const { return oh_; }
    // Produces the next 8-bit sRGB image; `result` comes back EMPTY (return 0) once a temporal-area sequence is exhausted,
    // like the reference's result.empty() (:191).
    int nextFrame(std::vector<unsigned char>& result)
    {
        if (!src_) return MFSR_E_STATE;
        if (seq_.empty() || radius_ < 0) {               // pull the source once (temporal mode) / once per burst
            src_->reset();
            seq_.clear();
            pos_ = 0;
            while (const RawFrame* f = src_->nextFrame()) {
                if (!seq_.empty() && (f->w != seq_[0]->w || f->h != seq_[0]->h || f->gray != seq_[0]->gray)) return MFSR_E_INVALID;
                seq_.push_back(f);
            }
            if (seq_.empty()) return MFSR_E_INVALID;
        }
        const RawFrame* first = seq_[0];
        const size_t n = seq_.size();
        size_t lo = 0, hi = n, ref = 0;
        if (radius_ >= 0) {
            if (pos_ >= n) { result.clear(); return 0; }
            lo = pos_ > (size_t)radius_ ? pos_ - radius_ : 0;
            hi = std::min(n, pos_ + radius_ + 1);
            ref = pos_ - lo;
            pos_++;
        }
        std::vector<const void*> ptrs;
        for (size_t i = lo; i < hi; i++) ptrs.push_back(seq_[i]->px.data());
        if (!h_) {
            mfsr_params p;
            mfsr_default_params(&p);
            p.scale = scale_; p.lk_iterations = iterations_; p.merge_flags = MFSR_MERGE_GAMMA;
            while (p.levels > 1 && (std::min(first->w, first->h) >> (p.levels - 1)) < 2 * p.max_shift + p.tile_size) p.levels--;
            const int cap = radius_ >= 0 ? 2 * radius_ + 1 : (int)n;
            const int rc = mfsr_create(&p, 0, first->w, first->h, cap, &h_);
            if (rc) return rc;
            mfsr_output_size(h_, first->w, first->h, &ow_, &oh_);
        }
        int rc = mfsr_set_frames(h_, ptrs.data(), (int)ptrs.size(), first->w, first->h, (int64_t)first->w * 2,
                                 first->gray ? MFSR_FMT_GRAY_U16 : MFSR_FMT_BAYER_U16, (int)ref, /*on_host*/1);
        if (rc) return rc;
        result.resize((size_t)ow_ * oh_ * 3);
        return mfsr_run_format(h_, result.data(), (int64_t)ow_ * 3, /*out_on_host*/1, MFSR_OUT_U8, /*async*/0);
    }
private:
    mfsr_handle h_ = nullptr;
    BurstFrameSource* src_ = nullptr;
    std::vector<const RawFrame*> seq_;
    size_t pos_ = 0;
    int scale_ = 2, iterations_ = 3, radius_ = -1, ow_ = 0, oh_ = 0;
};

void usage()
{
    printf("./multi_frame_sr_b200 [--radius R] optFlowName inputName iterations [frames]\n");
    printf("\toptFlowName: farneback, tvl1, brox, pyrlk (accepted for compatibility)\n");
    printf("\tinputName: city, car, iso, or a printf pattern of .pgm/.ppm frames with [frames]\n");
    printf("\titerations: integer, 1, 10, etc.\n");
    printf("\t--radius R: temporal-area mode (the reference uses 1): one result per frame from frames [i-R, i+R]\n");
}

}  // namespace

int main(int argc, char** argv)
{
    std::string flow = "farneback", input = "city";
    int iterations = 10, n_override = 0, radius = -1;
    for (int i = 1; i + 1 < argc; i++) {
        if (strcmp(argv[i], "--radius") == 0) {
            radius = atoi(argv[i + 1]);
            for (int j = i; j + 2 < argc; j++) argv[j] = argv[j + 2];
            argc -= 2;
            break;
        }
    }
    if (argc == 4 || argc == 5) {
        flow = argv[1]; input = argv[2]; iterations = atoi(argv[3]);
        if (iterations < 1) iterations = 1;
        if (argc == 5) n_override = atoi(argv[4]);
    } else if (argc != 1) {
        usage();
        return -1;
    }
    const int scale = 2, num_times = 10, real_times = 5;
    int num_images;
    std::string pattern, tag = input;
    if (input == "city") { num_images = 5; pattern = "img_%06d"; }
    else if (input == "car") { num_images = 4; pattern = "car/%d"; }
    else if (input == "iso") { num_images = 4; pattern = "iso/%06d"; }
    else if (n_override > 0) { num_images = n_override; pattern = input; tag = "burst"; }
    else { printf("wrong input\n"); return -1; }

    std::vector<RawFrame> frames;
    char buf[4096];
    for (int base = 1; base >= 0 && frames.empty(); base--) {
        for (int i = 0; i < num_images; i++) {
            RawFrame fr;
            bool ok = false;
            for (const char* ext : {"", ".pgm", ".ppm"}) {
                snprintf(buf, sizeof buf, (pattern + ext).c_str(), i + base);
                if ((ok = load_frame(buf, fr))) break;
            }
            if (!ok) { frames.clear(); break; }
            printf("%s, [%d x %d]\n", buf, fr.w, fr.h);
            frames.push_back(std::move(fr));
        }
    }
    if (frames.empty()) {
        snprintf(buf, sizeof buf, pattern.c_str(), 1);
        printf("cannot read %s(.pgm|.ppm)\n", buf);
        return -1;
    }
    BurstFrameSource source(std::move(frames), radius >= 0 ? num_times : 1);
    BurstSuperResolution sr;
    sr.setScale(scale);
    sr.setIterations(iterations);
    sr.setTemporalAreaRadius(radius);
    sr.setInput(&source);
    std::vector<unsigned char> result, last;
    std::chrono::steady_clock::time_point t0;
    double produced = 0;                       // results inside the timed part
    if (radius >= 0) {
        const int start_i = (num_times - real_times) * num_images;
        for (int i = 0; i < num_images * num_times; i++) {
            if (i == start_i) t0 = std::chrono::steady_clock::now();
            const int rc = sr.nextFrame(result);
            if (rc) { fprintf(stderr, "mfsr: %s (%d)\n", mfsr_error_string(rc), rc); return 1; }
            if (result.empty()) break;
            if (i >= start_i) produced += 1;
            last.swap(result);
        }
        result.swap(last);
    } else {
        for (int t = 0; t < num_times; t++) {
            if (t == num_times - real_times) t0 = std::chrono::steady_clock::now();
            const int rc = sr.nextFrame(result);
            if (rc) { fprintf(stderr, "mfsr: %s (%d)\n", mfsr_error_string(rc), rc); return 1; }
        }
        produced = real_times;
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("%g sec\n", sec);
    printf("%g FPS\n", (double)(real_times * num_images) / sec);
    const int ow = sr.outWidth(), oh = sr.outHeight();
    if (!write_ppm(tag + "_" + flow + "_sr_result.ppm", result, ow, oh) ||
        !write_ppm(tag + "_" + flow + "_sr2_result.ppm", sharpen(result, ow, oh), ow, oh)) {
        fprintf(stderr, "cannot write the result images\n");
        return 1;
    }
    printf("{\"output_megapixels_per_second\": %.2f, \"frames\": %d, \"out\": [%d, %d], \"scale\": %d, \"lk_iterations\": %d, "
           "\"temporal_radius\": %d}\n",
           produced * (double)ow * oh / 1e6 / sec, radius >= 0 ? std::min(num_images * num_times, 2 * radius + 1) : num_images, ow, oh, scale,
           iterations, radius);
    return 0;
}
