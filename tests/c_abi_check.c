/* Compiled as plain C99 against include/fri_cuda.h and linked with libfri_cuda.so by
 * tests/test_host.py::test_header_is_valid_c_and_links: the header must be a C header, every
 * prototype must resolve, and the host-only paths must behave without a device.  What a cgo /
 * JNI / Rust-bindgen consumer of the ABI would see. */
#include <stdio.h>
#include <string.h>

#include "fri_cuda.h"

typedef void (*fn)(void);

int main(void)
{
    /* every entry point of the header, by address (forces the linker to resolve all of them) */
    fn all[] = {
        (fn)fri_version, (fn)fri_last_error, (fn)fri_device_count, (fn)fri_plan_create, (fn)fri_plan_destroy,
        (fn)fri_plan_num_tiles, (fn)fri_plan_num_built, (fn)fri_plan_num_full_tiles, (fn)fri_plan_coefs_per_frame,
        (fn)fri_plan_pixels_covered, (fn)fri_plan_launch_info, (fn)fri_plan_centers, (fn)fri_plan_masks,
        (fn)fri_encode_tq_device, (fn)fri_decode_tq_device, (fn)fri_encode_tq_device16, (fn)fri_decode_tq_device16,
        (fn)fri_encode_tq, (fn)fri_decode_tq, (fn)fri_encode_tq16, (fn)fri_decode_tq16,
        (fn)fri_plan_emission_count, (fn)fri_plan_emission_order, (fn)fri_emit_device, (fn)fri_encode_tq_emit,
        (fn)fri_emit_device16, (fn)fri_encode_tq_emit16, (fn)fri_unemit_device, (fn)fri_unemit_device16,
        (fn)fri_decode_tq_emit, (fn)fri_decode_tq_emit16, (fn)fri_plan_emission_packed_bytes, (fn)fri_plan_emission_packed_size, (fn)fri_emit_device_packed,
        (fn)fri_encode_tq_emit_packed, (fn)fri_unemit_device_packed, (fn)fri_decode_tq_emit_packed, (fn)fri_emit_device10,
        (fn)fri_encode_tq_emit10, (fn)fri_unemit_device10, (fn)fri_decode_tq_emit10, (fn)fri_predict_device, (fn)fri_fit_parameters, (fn)fri_fit_device, (fn)fri_plan_part, (fn)fri_encode_tq_device_part, (fn)fri_decode_tq_device_part, (fn)fri_plan_groups_in_rows, (fn)fri_decode_tq_device_groups, (fn)fri_predict_host, (fn)fri_frv_pack, (fn)fri_frv_unpack,
        (fn)fri_frv_info, (fn)fri_frv_encode, (fn)fri_frv_decode, (fn)fri_frv_free, (fn)fri_plan_set_bands, (fn)fri_plan_set_async, (fn)fri_plan_set_independent_calls, (fn)fri_plan_sync, (fn)fri_host_alloc, (fn)fri_host_free,
        (fn)fri_plan_last_launches, (fn)fri_quant_divide, (fn)fri_quant_divide_magic, (fn)fri_quant_divide_small,
    };
    size_t i, n = sizeof(all) / sizeof(all[0]);
    for (i = 0; i < n; ++i)
        if (!all[i]) return 10;
    if (!strstr(fri_version(), "sm_100a")) return 11;

    /* host-only plan (device = -1): metadata works, compute refuses — there is no CPU fallback */
    fri_plan *plan = NULL;
    if (fri_plan_create(&plan, -1, 512, 512, 1, FRI_BASE_DEPTH, 1) != FRI_OK) return 12;
    if (fri_plan_num_built(plan) != 617 || fri_plan_num_tiles(plan) != 578 || fri_plan_num_full_tiles(plan) != 448) return 13;
    if (fri_plan_coefs_per_frame(plan) != 578u * 512u || fri_plan_pixels_covered(plan) != 512u * 512u) return 14;
    {
        static unsigned char px[512 * 512];
        static int32_t coefs[578 * 512];
        int rc = fri_encode_tq(plan, px, 1, NULL, coefs);
        if (rc != FRI_E_CUDA || !strstr(fri_last_error(), "no CPU fallback")) return 15;
    }
    if (fri_plan_set_bands(plan, 99) != FRI_E_INVALID) return 16;
    if (fri_quant_divide(-7, 4) != -1 || fri_quant_divide_small(-255, 5) != -51) return 17;  /* truncation toward zero */
    fri_plan_destroy(plan);
    printf("c abi ok: %u entry points, %s\n", (unsigned)n, fri_version());
    return 0;
}
