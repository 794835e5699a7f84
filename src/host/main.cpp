// main.cpp -- the pics.txt-driven host flow of the reference's stereo_matching/main.cpp, for the
// ASW hot path only, on top of the C ABI (include/asw_b200.h).  What is kept from the reference:
//   * pics.txt: pairs of whitespace-separated tokens, left then right (main.cpp:134-148);
//     the output folder is the path prefix before the first '/' (main.cpp:151-156)
//   * per image: decode both PNGs to RGBA8 (main.cpp:183-186), 10 runs (main.cpp:213),
//     PNG written from the first run, per-run timing row in a TSV log named after the device
//     (main.cpp:164-166,179-181,634-708) with the reference's column names
//   * per-stage timings come from device events (CL profiling there, CUDA events here)
// What is replaced: the OpenCL platform/device/context/program/queue setup and the enqueue
// sequence (main.cpp:119-130,158-172,210-256,434-526) -> asw_create + asw_disparity.
// --method hot (default) runs the hot path only and writes asw_disparity<suffix>.png = the WTA image;
// --method whole runs the whole ASW method (asw_stereo: + consistency, k refinement rounds, median,
// main.cpp:529-631) and writes the reference's three ASW PNGs (main.cpp:621-631) with the suffix;
// --method cross runs the cross-based method (asw_cross_stereo, main.cpp:258-367) and writes
// cross_based_initial / cross_based_disparity / median<suffix>.png (main.cpp:357-367);
// --method both = cross + whole, the reference's full per-run sequence.  Columns of a method that did
// not run are written as 0.
//
// --devices 0,1,2,3 (with --method hot) splits every frame into row bands over the listed GPUs (asw_multi_disparity:
// the reference visits its devices one after the other with the whole job, main.cpp:158-172; here they share a frame).
//
// Usage: stereo_matching [--pics pics.txt] [--root DIR] [--runs 10] [--device 0 | --devices 0,1,...] [--ndisp 61]
//                        [--iterations 7] [--method hot|whole|cross|both] [--refine 6] [--out-suffix _wta] [--log FILE]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <utility>
#include <vector>

#include "asw_b200.h"
#include "png_io.h"

struct Image {
    std::vector<unsigned char> pixel;
    unsigned width = 0, height = 0;
};

static const char* kHeader =
    "id\tmedL_solo\tmedR_solo\tmed_full\tcross_h\tcross_v\tcross_full\taggregation\tintegral_h\taggr_h\tintegral_v\taggr_v\t"
    "init_disp\tfinal_disp\tcross method total\t\t\taggr\tsupp_w\tv_aggr_mean\th_aggr_mean\ttotal aggregation\twta\t"
    "consistency\tv_ref_mean_L\tv_ref_mean_R\th_ref_mean_L\th_ref_mean_R\twta_mean_LR\tconsistency_mean\ttotal refinement\t"
    "median\ttotal WTA method";   // main.cpp:181

int main(int argc, char** argv) {
    std::string pics = "pics.txt", root = ".", suffix = "_wta", log_name;
    int runs = 10, device = 0, refine = 6;   // k = 6, main.cpp:176
    std::vector<int> devices;
    bool whole = false, hot = true, cross = false;
    asw_params prm;
    asw_params_default(&prm);
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&](const char* what) -> const char* {
            if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", what); exit(2); }
            return argv[++i];
        };
        if (a == "--pics") pics = next("--pics");
        else if (a == "--root") root = next("--root");
        else if (a == "--runs") runs = atoi(next("--runs"));
        else if (a == "--device") device = atoi(next("--device"));
        else if (a == "--devices") {
            std::string v = next("--devices");
            for (size_t p = 0; p < v.size();) {
                size_t q = v.find(',', p);
                if (q == std::string::npos) q = v.size();
                devices.push_back(atoi(v.substr(p, q - p).c_str()));
                p = q + 1;
            }
            if (!devices.empty()) device = devices[0];
        }
        else if (a == "--ndisp") prm.ndisp = atoi(next("--ndisp"));
        else if (a == "--iterations") prm.iterations = atoi(next("--iterations"));
        else if (a == "--refine") refine = atoi(next("--refine"));
        else if (a == "--method") {
            std::string m = next("--method");
            if (m != "hot" && m != "whole" && m != "cross" && m != "both") { fprintf(stderr, "--method must be hot, whole, cross or both\n"); return 2; }
            whole = m == "whole" || m == "both";
            cross = m == "cross" || m == "both";
            hot = m == "hot";
        }
        else if (a == "--out-suffix") suffix = next("--out-suffix");
        else if (a == "--log") log_name = next("--log");
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }

    // read image paths: pairs of tokens (main.cpp:134-148)
    std::vector<std::string> left_path, right_path, folder_name;
    {
        std::ifstream in(pics);
        if (!in) { fprintf(stderr, "cannot open %s\n", pics.c_str()); return 1; }
        std::string l, r;
        while (in >> l >> r) {
            left_path.push_back(l);
            right_path.push_back(r);
            folder_name.push_back(l.substr(0, l.find('/')));   // main.cpp:151-156
        }
    }

    asw_ctx* ctx = nullptr;
    int st = asw_create(&ctx, device);
    if (st != ASW_OK) {
        fprintf(stderr, "asw_create(device %d) failed: %s (a CUDA device is required)\n", device, asw_strerror(st));
        return 1;
    }
    asw_multi* multi = nullptr;
    if (devices.size() > 1) {
        if (!hot) { fprintf(stderr, "--devices needs --method hot (the frame split covers the hot path)\n"); asw_destroy(ctx); return 2; }
        st = asw_multi_create(&multi, devices.data(), (int)devices.size());
        if (st != ASW_OK) { fprintf(stderr, "asw_multi_create failed: %s\n", asw_strerror(st)); asw_destroy(ctx); return 1; }
        printf("\t- %zu devices share every frame (row bands, halo rows exchanged between iterations)\n", devices.size());
    }
    char dev_name[256] = "";
    asw_device_info(ctx, nullptr, nullptr, nullptr, dev_name, sizeof dev_name);
    printf("\t- Device name: %s\n", dev_name);
    if (log_name.empty()) log_name = root + "/" + dev_name;     // log named after the device (main.cpp:164-166)
    FILE* to_file = fopen(log_name.c_str(), "w");
    if (!to_file) { fprintf(stderr, "cannot open log %s\n", log_name.c_str()); asw_destroy(ctx); return 1; }

    int rc = 0;
    for (size_t img = 0; img < left_path.size(); img++) {
        fprintf(to_file, "\n%s - %s\n", dev_name, folder_name[img].c_str());
        printf("\n%s\n", folder_name[img].c_str());
        fprintf(to_file, "%s", kHeader);
        Image imgL, imgR;
        unsigned e1 = png_io::decode(imgL.pixel, imgL.width, imgL.height, root + "/" + left_path[img]);
        unsigned e2 = png_io::decode(imgR.pixel, imgR.width, imgR.height, root + "/" + right_path[img]);
        if (e1 || e2 || imgL.width != imgR.width || imgL.height != imgR.height) {
            fprintf(stderr, "cannot load pair %s / %s: %s\n", left_path[img].c_str(), right_path[img].c_str(),
                    png_io::error_text(e1 ? e1 : e2));
            rc = 1;
            continue;
        }
        const unsigned W = imgL.width, H = imgL.height;
        std::vector<unsigned char> disp((size_t)W * H * 4), pre, post;
        if (whole) { pre.resize(disp.size()); post.resize(disp.size()); }
        double sum_total = 0;
        for (int run = 0; run < runs; run++) {
            fprintf(to_file, "\nRun %d \t", run + 1);
            printf("\n---Working...\nRaw cost aggregation..  \ngestalt principle - support area.. \nCost aggregation.. \nWTA.. ");
            asw_timing t;
            asw_tail_timing tt;
            asw_cross_timing ct;
            memset(&t, 0, sizeof t);
            memset(&tt, 0, sizeof tt);
            memset(&ct, 0, sizeof ct);
            if (cross) {                                                             // main.cpp:258-367
                std::vector<unsigned char> ini(disp.size()), fin(disp.size()), med(disp.size());
                asw_cross_params cprm;
                asw_cross_params_default(&cprm);
                cprm.ndisp = prm.ndisp;
                st = asw_cross_stereo(ctx, imgL.pixel.data(), imgR.pixel.data(), (int)W, (int)H, &cprm, ini.data(), fin.data(), med.data(), &ct);
                if (st != ASW_OK) {
                    printf("error executing the cross-based method: %d (%s)\n", st, asw_last_error(ctx));
                    rc = 1;
                } else if (run == 0) {
                    const std::string dir = root + "/" + folder_name[img] + "/";
                    const std::pair<const char*, const std::vector<unsigned char>*> outs[3] = {
                        {"cross_based_initial", &ini}, {"cross_based_disparity", &fin}, {"median", &med}};
                    for (const auto& o : outs) {
                        unsigned e = png_io::encode(dir + o.first + suffix + ".png", *o.second, W, H);
                        if (e) { fprintf(stderr, "cannot write %s%s%s.png: %s\n", dir.c_str(), o.first, suffix.c_str(), png_io::error_text(e)); rc = 1; }
                    }
                }
            }
            if (!whole && !hot) st = ASW_OK;
            else if (whole)
                st = asw_stereo(ctx, imgL.pixel.data(), imgR.pixel.data(), (int)W, (int)H, &prm, refine, disp.data(), pre.data(), post.data(), &t, &tt);
            else if (multi) {
                asw_multi_timing mt;
                st = asw_multi_disparity(multi, imgL.pixel.data(), imgR.pixel.data(), (int)W, (int)H, &prm, disp.data(), nullptr, nullptr, &mt);
                t.agg_total_ms = t.total_ms = mt.compute_ms;   // the per-stage columns belong to one device; the frame total is what all share
                t.h2d_ms = mt.upload_ms;
                t.d2h_ms = mt.download_ms;
            } else
                st = asw_disparity(ctx, imgL.pixel.data(), imgR.pixel.data(), (int)W, (int)H, &prm, disp.data(), nullptr, nullptr, &t);
            if (st != ASW_OK) {
                printf("ASW error executing hot path: %d (%s)\n", st, multi ? asw_multi_last_error(multi) : asw_last_error(ctx));   // ErCheck prints and continues
                rc = 1;
                continue;
            }
            if (run == 0 && (whole || hot)) {   // the reference's PNGs on disk come from run 1
                const std::string dir = root + "/" + folder_name[img] + "/";
                auto save = [&](const std::string& name, const std::vector<unsigned char>& px) {
                    unsigned e = png_io::encode(dir + name + suffix + ".png", px, W, H);
                    if (e) { fprintf(stderr, "cannot write %s: %s\n", (dir + name + suffix + ".png").c_str(), png_io::error_text(e)); rc = 1; }
                };
                save("asw_disparity", disp);                                    // main.cpp:621-623
                if (whole) {
                    save("asw_consistency_pre-reff", pre);                      // main.cpp:625-627
                    save("asw_consistency_post-reff", post);                    // main.cpp:629-631
                }
            }
            // medL_solo medR_solo med_full cross_h cross_v cross_full aggregation integral_h aggr_h integral_v aggr_v
            // init_disp final_disp "cross method total" (main.cpp:379-396)
            fprintf(to_file, "%0.3f\t%0.3f\t%0.3f\t%0.3f\t%0.3f\t%0.3f\t%0.3f\t%0.3f\t%0.3f\t%0.3f\t%0.3f\t%0.3f\t%0.3f\t%0.3f\t", ct.median_l_ms,
                    ct.median_r_ms, ct.median_ms, ct.cross_l_ms, ct.cross_r_ms, ct.cross_ms, ct.aggregation_ms, ct.integral_h_ms, ct.oii_h_ms,
                    ct.integral_v_ms, ct.oii_v_ms, ct.init_disparity_ms, ct.final_disparity_ms, ct.total_ms);
            fprintf(to_file, "\t\t");
            fprintf(to_file, "%0.3f\t%0.3f\t%0.3f\t%0.3f\t%0.3f\t%0.3f\t", t.raw_ms, t.supp_ms, t.vagg_mean_ms, t.hagg_mean_ms,
                    t.agg_total_ms, t.wta_ms);
            // consistency, v_ref_mean_L/R, h_ref_mean_L/R, wta_mean_LR, consistency_mean, total refinement, median
            // (main.cpp:661-708; zero with --method hot); the right-view WTA is folded into the first consistency column
            fprintf(to_file, "%0.3f\t%0.3f\t%0.3f\t%0.3f\t%0.3f\t%0.3f\t%0.3f\t%0.3f\t%0.3f\t", tt.right_wta_ms + tt.consistency_ms,
                    tt.vref_mean_l_ms, tt.vref_mean_r_ms, tt.href_mean_l_ms, tt.href_mean_r_ms, tt.wta_ref_mean_ms, tt.consistency_mean_ms,
                    tt.refinement_total_ms, tt.median_ms);
            const float whole_ms = whole ? tt.total_ms : hot ? t.total_ms : ct.total_ms;
            fprintf(to_file, "%0.3f\t", whole_ms);
            sum_total += whole_ms;
        }
        if (runs > 0) {
            double ms = sum_total / runs;
            printf("\n%s: %ux%u, %d disparities, r = %d: %.3f ms/frame (device), %.1f Mpix*disp/s\n", folder_name[img].c_str(), W, H,
                   prm.ndisp, prm.iterations, ms, ms > 0 ? (double)W * H * prm.ndisp / ms / 1e3 : 0.0);
        }
    }
    fclose(to_file);
    if (multi) asw_multi_destroy(multi);
    asw_destroy(ctx);
    return rc;
}
