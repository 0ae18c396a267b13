// Headless file front end over include/stabilizer.hpp: the reference's main loop for --file input
// (/root/reference/src/main.cpp:196-236, src/main_utils.cpp:397-417 captureFrame, :459-493
// processAndDisplayFrames) without the GUI: frames come from a raw BGR24 stream (what
// `ffmpeg -f rawvideo -pix_fmt bgr24` writes) instead of cv::VideoCapture, the stabilized frames go
// to a raw BGR24 stream instead of imshow.  Flags follow the reference's CLI
// (src/main_utils.cpp:35-236: --file, --past-window, --future-window, --working-height); the
// stabilization-mode keys of handleStabilizationControls (:371-395) become --mode / --mode-at.
//
//   vstab_file --file in.bgr --width 1920 --height 1080 --out out.bgr
//              [--past-window 60] [--future-window 45] [--working-height 360]
//              [--mode global|lock|orb|sift|translation|rotation] [--mode-at CALL] [--side-by-side] [--frames N]
//
// --side-by-side writes [delayed original | stabilized] (2*width columns): the original is delayed by
// `future-window` frames through the same deque the reference keeps (originalFrameBuffer, :466-472), so both halves
// show the same source frame; nothing is written while that buffer fills (the reference prints "Buffering frames").
// "-" means stdin / stdout.  Exit status: 0 ok, 1 bad usage, 2 I/O, 3 stabilizer error.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "stabilizer.hpp"

namespace {

struct Options {
    std::string in, out = "-";
    int width = 0, height = 0;
    long frames = -1, mode_at = 0;
    long past = -1, future = -1;         // default: 2.0 s / 1.5 s at 30 fps (src/main.cpp:205-206)
    int working_height = 360;            // src/main_utils.hpp:27
    bool side_by_side = false, have_mode = false;
    StabilizationMode mode = StabilizationMode::GLOBAL_SMOOTHING;
};

bool parse_mode(const std::string& s, StabilizationMode& m) {
    if (s == "global") m = StabilizationMode::GLOBAL_SMOOTHING;
    else if (s == "lock") m = StabilizationMode::ACCUMULATED_FULL_LOCK;
    else if (s == "orb") m = StabilizationMode::ORB_FULL_LOCK;
    else if (s == "sift") m = StabilizationMode::SIFT_FULL_LOCK;
    else if (s == "translation") m = StabilizationMode::TRANSLATION_LOCK;
    else if (s == "rotation") m = StabilizationMode::ROTATION_LOCK;
    else return false;
    return true;
}

int usage(const char* argv0) {
    std::fprintf(stderr,
                 "usage: %s --file <in.bgr|-> --width W --height H [--out <out.bgr|->] [--past-window N] [--future-window N]\n"
                 "          [--working-height N] [--mode global|lock|orb|sift|translation|rotation] [--mode-at CALL]\n"
                 "          [--side-by-side] [--frames N]\n", argv0);
    return 1;
}

bool read_full(std::FILE* f, uint8_t* p, size_t n) {
    while (n) {
        const size_t r = std::fread(p, 1, n, f);
        if (r == 0) return false;
        p += r; n -= r;
    }
    return true;
}

}  // namespace

int main(int argc, char** argv) {
    Options o;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&](const char* what) -> const char* {
            if (i + 1 >= argc) { std::fprintf(stderr, "Error: %s requires a value.\n", what); std::exit(1); }
            return argv[++i];
        };
        if (a == "--help" || a == "-h") return usage(argv[0]);
        else if (a == "--file") o.in = next("--file");
        else if (a == "--out") o.out = next("--out");
        else if (a == "--width") o.width = std::atoi(next("--width"));
        else if (a == "--height") o.height = std::atoi(next("--height"));
        else if (a == "--frames") o.frames = std::atol(next("--frames"));
        else if (a == "--past-window") o.past = std::atol(next("--past-window"));
        else if (a == "--future-window") o.future = std::atol(next("--future-window"));
        else if (a == "--working-height") o.working_height = std::atoi(next("--working-height"));
        else if (a == "--mode") { if (!parse_mode(next("--mode"), o.mode)) return usage(argv[0]); o.have_mode = true; }
        else if (a == "--mode-at") o.mode_at = std::atol(next("--mode-at"));
        else if (a == "--side-by-side") o.side_by_side = true;
        else { std::fprintf(stderr, "Error: unknown argument %s\n", a.c_str()); return usage(argv[0]); }
    }
    if (o.in.empty() || o.width <= 0 || o.height <= 0) return usage(argv[0]);
    const double fps = 30.0;
    if (o.past < 0) o.past = (long)(2.0 * fps);
    if (o.future < 0) o.future = (long)(1.5 * fps);

    std::FILE* fin = o.in == "-" ? stdin : std::fopen(o.in.c_str(), "rb");
    std::FILE* fout = o.out == "-" ? stdout : std::fopen(o.out.c_str(), "wb");
    if (!fin || !fout) { std::fprintf(stderr, "Error: cannot open %s\n", !fin ? o.in.c_str() : o.out.c_str()); return 2; }

    const size_t step = (size_t)o.width * 3, nbytes = step * (size_t)o.height;
    uint8_t *hin = nullptr, *hout = nullptr;
    std::deque<std::vector<uint8_t>> originalFrameBuffer;       // src/main_utils.cpp:466
    std::vector<uint8_t> row;
    if (o.side_by_side) row.resize(2 * step);
    long n = 0, written = 0;
    int rc = 0;
    try {
        Stabilizer stabilizer((size_t)o.past, (size_t)o.future, o.working_height);   // argument errors throw before any device work
        // pinned staging buffers: the per-frame call copies straight from / into them
        hin = static_cast<uint8_t*>(vstab_host_alloc(nbytes));
        hout = static_cast<uint8_t*>(vstab_host_alloc(nbytes));
        if (!hin || !hout) throw std::runtime_error("pinned host allocation failed");
        const auto t0 = std::chrono::steady_clock::now();
        while ((o.frames < 0 || n < o.frames) && read_full(fin, hin, nbytes)) {
            if (o.have_mode && n == o.mode_at) stabilizer.setStabilizationMode(o.mode);
            stabilizer.stabilizeFrame(ImageView{hin, o.height, o.width, step}, ImageView{hout, o.height, o.width, step});
            ++n;
            if (!o.side_by_side) {
                if (std::fwrite(hout, 1, nbytes, fout) != nbytes) { rc = 2; break; }
                ++written;
                continue;
            }
            originalFrameBuffer.emplace_back(hin, hin + nbytes);
            if (originalFrameBuffer.size() > (size_t)o.future) {
                const std::vector<uint8_t>& delayed = originalFrameBuffer.front();
                for (int y = 0; y < o.height && rc == 0; ++y) {
                    std::memcpy(row.data(), delayed.data() + (size_t)y * step, step);
                    std::memcpy(row.data() + step, hout + (size_t)y * step, step);
                    if (std::fwrite(row.data(), 1, row.size(), fout) != row.size()) rc = 2;
                }
                originalFrameBuffer.pop_front();
                ++written;
                if (rc) break;
            }
        }
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::fprintf(stderr, "%ld frames in, %ld frames out, %.1f frames/s (window %zu)\n", n, written,
                     sec > 0 ? n / sec : 0.0, stabilizer.totalFrameWindowSize());
    } catch (const std::invalid_argument& e) {
        std::fprintf(stderr, "Error: %s\n", e.what());
        rc = 1;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "Error: %s\n", e.what());
        rc = 3;
    }
    if (hin) vstab_host_free(hin);
    if (hout) vstab_host_free(hout);
    if (fin != stdin) std::fclose(fin);
    if (fout != stdout) std::fclose(fout); else std::fflush(stdout);
    return rc;
}
