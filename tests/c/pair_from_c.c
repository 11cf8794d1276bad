/* A plain-C99 consumer of the C ABI (include/jsfe.h): what a cgo / JNI / N-API binding compiles against.
 * usage: pair_from_c <height> <width> <levels> <tile> <raw file with L then R image, h*w bytes each> <mb> <mbf>
 * prints: n_left n_right n_depth checksum(keypoint planes of the left eye) */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "jsfe.h"

int main(int argc, char** argv) {
    if (argc != 8) { fprintf(stderr, "usage: %s h w levels tile raw mb mbf\n", argv[0]); return 2; }
    jsfe_config cfg = {0};
    cfg.height = atoi(argv[1]); cfg.width = atoi(argv[2]); cfg.n_levels = atoi(argv[3]);
    cfg.scale_factor = 1.2f; cfg.fast_n_min = 9; cfg.fast_n_max = 14; cfg.th_fast_min = 7; cfg.th_fast_max = 20;
    cfg.tile_h = cfg.tile_w = atoi(argv[4]);
    cfg.nms_ms_mode_gpu = 1; cfg.max_images = 2;
    const size_t bytes = (size_t)cfg.height * cfg.width * 2;
    uint8_t* images = (uint8_t*)malloc(bytes);
    FILE* f = fopen(argv[5], "rb");
    if (!images || !f || fread(images, 1, bytes, f) != bytes) { fprintf(stderr, "cannot read %s\n", argv[5]); return 2; }
    fclose(f);
    jsfe_handle* h = NULL;
    if (jsfe_create(&cfg, &h) != JSFE_OK) { fprintf(stderr, "jsfe_create: %s\n", jsfe_last_error()); return 1; }
    jsfe_host_results r;
    if (jsfe_process_host_pairs(h, 1, images, 0, 100, 50, (float)atof(argv[6]), (float)atof(argv[7]), &r) != JSFE_OK) {
        fprintf(stderr, "jsfe_process_host_pairs: %s\n", jsfe_last_error());
        return 1;
    }
    const int nl = r.n_keypoints[0], nr = r.n_keypoints[1];
    int n_depth = 0;
    uint64_t sum = 0;
    for (int i = 0; i < nl; ++i) {
        if (r.depth[i] > 0.0f) ++n_depth;
        for (int pl = 0; pl < 6; ++pl) sum = sum * 1000003u + (uint32_t)r.kps[(size_t)pl * r.capacity + i];
    }
    printf("%d %d %d %llu\n", nl, nr, n_depth, (unsigned long long)sum);
    jsfe_destroy(h);
    free(images);
    return 0;
}
