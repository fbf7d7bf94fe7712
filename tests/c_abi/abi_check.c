/* Compiled and run by tests/test_host.py::test_header_is_plain_c_and_links: the header must be valid C99 for a caller that
 * knows nothing about CUDA or torch, and the library must resolve every entry point at link time.  No kernel is launched:
 * only the argument-checking paths run (they return before any CUDA call). */
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "orienmask_b200.h"

int main(void) {
    om_post_config cfg;
    size_t bytes = 0;
    int32_t rc;
    /* every entry point referenced so that the link fails if one is missing */
    typedef void (*fn)(void);
    fn entry[] = {(fn)om_abi_version, (fn)om_last_error, (fn)om_launch_count, (fn)om_launch_count_reset,
                  (fn)om_post_workspace_bytes, (fn)om_decode_select, (fn)om_batched_nms, (fn)om_mask_assemble,
                  (fn)om_nms, (fn)om_nms_ex, (fn)om_conv_create, (fn)om_conv_run, (fn)om_conv_run_to, (fn)om_conv_destroy,
                  (fn)om_stem_conv, (fn)om_preprocess, (fn)om_mask_rle, (fn)om_mask_areas, (fn)om_mask_blend,
                  (fn)om_debug_conv_plan_info, (fn)om_debug_conv_timeline, (fn)om_debug_trace, (fn)om_debug_phase_log, (fn)om_engine_workspace_bytes, (fn)om_engine_create,
                  (fn)om_forward, (fn)om_engine_destroy, (fn)om_engine_layer_count, (fn)om_engine_layer_info, (fn)om_engine_run_layer, (fn)om_engine_run_layers};
    memset(&cfg, 0, sizeof cfg);
    rc = om_post_workspace_bytes(&cfg, 1, &bytes);
    if (rc != OM_ERR_INVALID || strstr(om_last_error(), "num_scales") == NULL) return 1;
    if (om_nms(NULL, 2000, 0.5f, NULL, NULL, NULL) != OM_ERR_INVALID) return 2;
    if (om_conv_run(NULL, NULL) != OM_ERR_INVALID) return 3;
    om_conv_destroy(NULL);
    if (om_forward(NULL, NULL, NULL, NULL, NULL) != OM_ERR_INVALID) return 4;
    om_engine_destroy(NULL);
    printf("abi %d entries %d sizeof(om_post_config) %d sizeof(om_conv_desc) %d sizeof(om_prep_config) %d "
           "sizeof(om_rle_image) %d sizeof(om_blend_config) %d\n", (int)om_abi_version(), (int)(sizeof entry / sizeof entry[0]),
           (int)sizeof(om_post_config), (int)sizeof(om_conv_desc), (int)sizeof(om_prep_config), (int)sizeof(om_rle_image),
           (int)sizeof(om_blend_config));
    printf("offsetof(om_post_config.anchor_w) %d offsetof(om_post_config.nms_post) %d offsetof(om_conv_desc.input) %d "
           "offsetof(om_conv_desc.out_s2d) %d offsetof(om_prep_config.pad_value) %d offsetof(om_rle_image.vflip) %d "
           "offsetof(om_blend_config.alpha) %d\n", (int)offsetof(om_post_config, anchor_w), (int)offsetof(om_post_config, nms_post),
           (int)offsetof(om_conv_desc, input), (int)offsetof(om_conv_desc, out_s2d), (int)offsetof(om_prep_config, pad_value),
           (int)offsetof(om_rle_image, vflip), (int)offsetof(om_blend_config, alpha));
    return 0;
}
