// GPU JPEG encode for the traversal / sampling output stage (traverse_latent_space.py:26-41,466-490; sample_gan.py:172-176):
// the reference pulls every fp32 frame to the host, normalises it there and encodes with PIL (quality 95, optimised Huffman
// tables, progressive).  At config-5 rates (thousands of 1024^2 frames per second per GPU) the host encode is the bottleneck
// (SURVEY.md §8f rank 2), so the frames - already uint8 on the device (imgio.cu, bit-identical tensor2image) - are encoded by
// nvJPEG on the GPU with the same settings and only the compressed bitstream crosses PCIe.
//
// nvJPEG is a CUDA-toolkit library (like cuBLAS); it is resolved with dlopen at first use so that libwgs_b200.so loads - and
// every other entry point works - on a box without it.  No CPU fallback: without nvJPEG the call fails with a message.
#include "common.cuh"
#include "wgs_b200.h"
#include <dlfcn.h>
#include <nvjpeg.h>
#include <mutex>

namespace wgs {

struct NvJpegApi {
    void* lib = nullptr;
    bool tried = false;
    decltype(&nvjpegCreateSimple) create = nullptr;
    decltype(&nvjpegEncoderStateCreate) state_create = nullptr;
    decltype(&nvjpegEncoderParamsCreate) params_create = nullptr;
    decltype(&nvjpegEncoderParamsSetQuality) set_quality = nullptr;
    decltype(&nvjpegEncoderParamsSetEncoding) set_encoding = nullptr;
    decltype(&nvjpegEncoderParamsSetOptimizedHuffman) set_huffman = nullptr;
    decltype(&nvjpegEncoderParamsSetSamplingFactors) set_sampling = nullptr;
    decltype(&nvjpegEncodeImage) encode_image = nullptr;
    decltype(&nvjpegEncodeYUV) encode_yuv = nullptr;
    decltype(&nvjpegEncodeRetrieveBitstream) retrieve = nullptr;
    nvjpegHandle_t handle = nullptr;
    nvjpegEncoderState_t state = nullptr;
    nvjpegEncoderParams_t params = nullptr;
    int quality = -1, progressive = -1, gray = -1;
};

static NvJpegApi g_jpeg;
static std::mutex g_jpeg_mutex;

static bool jpeg_load(std::string& why) {
    NvJpegApi& a = g_jpeg;
    if (a.tried) { why = "nvJPEG is not available on this machine"; return a.lib != nullptr; }
    a.tried = true;
    const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12",
                           "/usr/local/cuda/targets/x86_64-linux/lib/libnvjpeg.so.12"};
    for (const char* n : names) {
        a.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (a.lib) break;
    }
    if (!a.lib) { why = "libnvjpeg.so.12 not found (dlopen)"; return false; }
#define WGS_JPEG_SYM(field, name) \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, #name)); \
    if (!a.field) { why = "nvJPEG symbol missing: " #name; dlclose(a.lib); a.lib = nullptr; return false; }
    WGS_JPEG_SYM(create, nvjpegCreateSimple)
    WGS_JPEG_SYM(state_create, nvjpegEncoderStateCreate)
    WGS_JPEG_SYM(params_create, nvjpegEncoderParamsCreate)
    WGS_JPEG_SYM(set_quality, nvjpegEncoderParamsSetQuality)
    WGS_JPEG_SYM(set_encoding, nvjpegEncoderParamsSetEncoding)
    WGS_JPEG_SYM(set_huffman, nvjpegEncoderParamsSetOptimizedHuffman)
    WGS_JPEG_SYM(set_sampling, nvjpegEncoderParamsSetSamplingFactors)
    WGS_JPEG_SYM(encode_image, nvjpegEncodeImage)
    WGS_JPEG_SYM(encode_yuv, nvjpegEncodeYUV)
    WGS_JPEG_SYM(retrieve, nvjpegEncodeRetrieveBitstream)
#undef WGS_JPEG_SYM
    return true;
}

}  // namespace wgs

using namespace wgs;

#define WGS_JPEG(expr)                                                                             \
    do { nvjpegStatus_t s__ = (expr);                                                              \
         if (s__ != NVJPEG_STATUS_SUCCESS)                                                         \
             return ::wgs::fail(__FILE__, __LINE__, std::string(#expr) + ": nvjpeg status " + std::to_string((int)s__)); \
    } while (0)

extern "C" int wgs_jpeg_available(void) {
    std::lock_guard<std::mutex> lock(g_jpeg_mutex);
    std::string why;
    return jpeg_load(why) ? 1 : 0;
}

extern "C" int wgs_jpeg_encode(const unsigned char* pixels, int N, int H, int W, int C, int quality, int progressive,
                               unsigned char* h_out, long long h_capacity_per_image, long long* h_sizes, void* stream) {
    WGS_REQUIRE(pixels && h_out && h_sizes && N >= 0 && H >= 1 && W >= 1, "jpeg_encode: bad arguments");
    WGS_REQUIRE(C == 3 || C == 1, "jpeg_encode: 3 (interleaved RGB) or 1 (gray) channels");
    WGS_REQUIRE(quality >= 1 && quality <= 100, "jpeg_encode: quality must be in 1..100");
    std::lock_guard<std::mutex> lock(g_jpeg_mutex);
    std::string why;
    WGS_REQUIRE(jpeg_load(why), "jpeg_encode: " + why + " (there is no CPU fallback in libwgs_b200)");
    NvJpegApi& a = g_jpeg;
    const cudaStream_t st = (cudaStream_t)stream;
    if (!a.handle) {
        WGS_JPEG(a.create(&a.handle));
        WGS_JPEG(a.state_create(a.handle, &a.state, st));
        WGS_JPEG(a.params_create(a.handle, &a.params, st));
    }
    if (a.quality != quality || a.progressive != progressive || a.gray != (C == 1)) {
        WGS_JPEG(a.set_quality(a.params, quality, st));
        WGS_JPEG(a.set_huffman(a.params, 1, st));                                   // PIL optimize=True
        WGS_JPEG(a.set_encoding(a.params, progressive ? NVJPEG_ENCODING_PROGRESSIVE_DCT_HUFFMAN : NVJPEG_ENCODING_BASELINE_DCT, st));
        // PIL's default chroma sub-sampling is 4:2:0 at every quality setting below "keep"
        WGS_JPEG(a.set_sampling(a.params, C == 1 ? NVJPEG_CSS_GRAY : NVJPEG_CSS_420, st));
        a.quality = quality; a.progressive = progressive; a.gray = (C == 1);
    }
    for (int i = 0; i < N; ++i) {
        nvjpegImage_t img;
        memset(&img, 0, sizeof(img));
        img.channel[0] = const_cast<unsigned char*>(pixels) + (size_t)i * H * W * C;
        img.pitch[0] = (size_t)W * C;
        if (C == 3) WGS_JPEG(a.encode_image(a.handle, a.state, a.params, &img, NVJPEG_INPUT_RGBI, W, H, st));
        else WGS_JPEG(a.encode_yuv(a.handle, a.state, a.params, &img, NVJPEG_CSS_GRAY, W, H, st));
        size_t len = 0;
        WGS_JPEG(a.retrieve(a.handle, a.state, nullptr, &len, st));
        WGS_REQUIRE((long long)len <= h_capacity_per_image, "jpeg_encode: output buffer too small for image " + std::to_string(i));
        WGS_JPEG(a.retrieve(a.handle, a.state, h_out + (size_t)i * h_capacity_per_image, &len, st));
        h_sizes[i] = (long long)len;
        count_launch();
    }
    WGS_CUDA(cudaStreamSynchronize(st));
    return 0;
}
