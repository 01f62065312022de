// vgaudio_batch — batch conversion of a directory of WAVE files to .dsp / .adx / .hca on the GPU.
//
// The counterpart of `VGAudioCli -b` (src/VGAudio.Cli/Batch.cs:11-51): the reference enumerates the input files and runs
// Convert.ConvertFile on each from a Parallel.ForEach; here the host only reads and writes files, and every
// WaveReader -> encoder -> writer chain of a chunk of files runs as ONE coalesced call on the device
// (vgb_convert_wave_batch).  A file that fails is reported and skipped, like the reference's try/catch (:39-43).
//
//   vgaudio_batch -i <indir> -o <outdir> --out-format dsp|adx|hca|wav [-r]   (wav: .dsp inputs are decoded) [--no-trim] [--hcaquality Highest|High|Middle|Low|Lowest]
//                 [--bitrate N] [--limit-bitrate] [--keycode N] [--keystring S] [--adxtype Linear|Fixed|Exp|ExpEnc...]
//                 [--framesize N] [--version 3|4] [--chunk-mb N]
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <string>
#include <vector>

#include "../../include/vgaudio_b200.h"

namespace fs = std::filesystem;

static bool read_file(const fs::path &p, std::vector<uint8_t> &out)
{
    std::ifstream f(p, std::ios::binary | std::ios::ate);
    if (!f) return false;
    const std::streamsize n = f.tellg();
    f.seekg(0);
    out.resize((size_t)n);
    return n == 0 || (bool)f.read(reinterpret_cast<char *>(out.data()), n);
}

static int usage()
{
    std::fprintf(stderr, "usage: vgaudio_batch -i <indir> -o <outdir> --out-format dsp|adx|hca|wav [-r] [--no-trim] [--hcaquality Q] [--bitrate N]\n"
                         "                     [--limit-bitrate] [--keycode N] [--keystring S] [--adxtype linear|fixed|exp] [--framesize N] [--version 3|4]\n"
                         "                     [--chunk-mb N]\n");
    return 2;
}

int main(int argc, char **argv)
{
    std::string in_dir, out_dir, fmt, key_string;
    bool recurse = false, have_code = false;
    uint64_t key_code = 0;
    size_t chunk_mb = 1024;
    vgb_convert_options opt{};
    opt.hca_key_type = -1;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "-i") in_dir = next();
        else if (a == "-o") out_dir = next();
        else if (a == "--out-format") fmt = next();
        else if (a == "-r") recurse = true;
        else if (a == "--no-trim") opt.no_trim = 1;
        else if (a == "--bitrate") opt.hca_bitrate = std::atoi(next());
        else if (a == "--limit-bitrate") opt.hca_limit_bitrate = 1;
        else if (a == "--keycode") { key_code = std::strtoull(next(), nullptr, 0); have_code = true; }
        else if (a == "--keystring") key_string = next();
        else if (a == "--framesize") opt.adx_frame_size = std::atoi(next());
        else if (a == "--version") opt.adx_version = std::atoi(next());
        else if (a == "--chunk-mb") chunk_mb = (size_t)std::atoll(next());
        else if (a == "--hcaquality") {
            const std::string q = next();
            const char *names[] = {"", "highest", "high", "middle", "low", "lowest"};
            for (int k = 1; k <= 5; k++) if (strcasecmp(q.c_str(), names[k]) == 0) opt.hca_quality = k;
            if (!opt.hca_quality) return usage();
        } else if (a == "--adxtype") {
            const std::string t = next();
            opt.adx_type = strcasecmp(t.c_str(), "fixed") == 0 ? 2 : strcasecmp(t.c_str(), "exp") == 0 ? 4 : 3;
        } else return usage();
    }
    if (in_dir.empty() || out_dir.empty()) return usage();
    const bool to_wave = fmt == "wav";  // the decode direction: .dsp files in, 16-bit WAVE files out
    if (fmt == "dsp") opt.out_type = VGB_CONTAINER_DSP;
    else if (fmt == "adx") opt.out_type = VGB_CONTAINER_ADX;
    else if (fmt == "hca") opt.out_type = VGB_CONTAINER_HCA;
    else if (!to_wave) return usage();
    if (opt.out_type == VGB_CONTAINER_ADX && (have_code || !key_string.empty())) {
        vgb_adx_key k{};
        const int32_t s = !key_string.empty() ? vgb_adx_key_from_string(key_string.c_str(), &k) : vgb_adx_key_from_code(key_code, &k);
        if (s != VGB_OK) { std::fprintf(stderr, "%s\n", vgb_last_error()); return 1; }
        opt.adx_has_key = 1; opt.adx_key_seed = k.seed; opt.adx_key_mult = k.mult; opt.adx_key_inc = k.inc;
        opt.adx_encryption_type = !key_string.empty() ? 8 : 9;  // CreateConfiguration.cs:126-135: key strings are type 8, key codes type 9
    }
    if (opt.out_type == VGB_CONTAINER_HCA && have_code) { opt.hca_key_type = 56; opt.hca_key_code = key_code; }

    // Batch.cs:16-19: the files of the input directory (here: the WAVE ones; the other containers are not read)
    std::vector<fs::path> files;
    std::error_code ec;
    auto take = [&](const fs::directory_entry &e) {
        if (!e.is_regular_file()) return;
        std::string ext = e.path().extension().string();
        std::transform(ext.begin(), ext.end(), ext.begin(), ::tolower);
        if (to_wave ? ext == ".dsp" : (ext == ".wav" || ext == ".wave")) files.push_back(e.path());
    };
    if (recurse) for (auto &e : fs::recursive_directory_iterator(in_dir, ec)) take(e);
    else for (auto &e : fs::directory_iterator(in_dir, ec)) take(e);
    if (ec) { std::fprintf(stderr, "cannot read %s: %s\n", in_dir.c_str(), ec.message().c_str()); return 1; }
    std::sort(files.begin(), files.end());
    if (vgb_init(0, 0) != VGB_OK) { std::fprintf(stderr, "%s\n", vgb_last_error()); return 1; }

    const auto t0 = std::chrono::steady_clock::now();
    size_t done = 0, failed = 0;
    uint64_t bytes_in = 0, bytes_out = 0;
    for (size_t first = 0; first < files.size();) {
        // a chunk of files that fits the host budget
        std::vector<std::vector<uint8_t>> in;
        size_t last = first, held = 0;
        while (last < files.size() && (last == first || held < (chunk_mb << 20))) {
            in.emplace_back();
            if (!read_file(files[last], in.back())) { std::fprintf(stderr, "Error reading %s\n", files[last].c_str()); in.back().clear(); }
            held += in.back().size();
            last++;
        }
        const int n = (int)(last - first);
        std::vector<const uint8_t *> ptr(n);
        std::vector<int64_t> len(n), out_size(n);
        std::vector<int32_t> status(n);
        for (int k = 0; k < n; k++) { ptr[k] = in[k].data(); len[k] = (int64_t)in[k].size(); bytes_in += in[k].size(); }
        auto convert = [&](uint8_t *const *outs) {
            return to_wave ? vgb_convert_dsp_to_wave_batch(ptr.data(), len.data(), n, out_size.data(), outs, status.data())
                           : vgb_convert_wave_batch(ptr.data(), len.data(), n, &opt, out_size.data(), outs, status.data(), nullptr, nullptr);
        };
        if (convert(nullptr) != VGB_OK) {
            std::fprintf(stderr, "%s\n", vgb_last_error());
            return 1;
        }
        std::vector<std::vector<uint8_t>> out(n);
        std::vector<uint8_t *> optr(n, nullptr);
        for (int k = 0; k < n; k++) if (status[k] == VGB_OK) { out[k].resize((size_t)out_size[k]); optr[k] = out[k].data(); }
        if (convert(optr.data()) != VGB_OK) {
            std::fprintf(stderr, "%s\n", vgb_last_error());
            return 1;
        }
        for (int k = 0; k < n; k++) {
            const fs::path &src = files[first + k];
            if (status[k] != VGB_OK) { std::fprintf(stderr, "Error converting %s\n", src.filename().c_str()); failed++; continue; }
            fs::path rel = fs::relative(src, in_dir, ec);
            fs::path dst = fs::path(out_dir) / rel;
            dst.replace_extension(fmt);                         // Path.ChangeExtension (Batch.cs:29)
            fs::create_directories(dst.parent_path(), ec);
            std::ofstream f(dst, std::ios::binary);
            f.write(reinterpret_cast<const char *>(out[k].data()), (std::streamsize)out[k].size());
            bytes_out += out[k].size();
            done++;
        }
        first = last;
    }
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("%zu files converted, %zu failed, %.1f MB in, %.1f MB out, %.3f s\n", done, failed, bytes_in / 1e6, bytes_out / 1e6, s);
    vgb_shutdown();
    return failed ? 3 : 0;
}
