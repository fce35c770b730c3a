/* CPU restatement of the reference's RoIAlign forward (TensorFlow-style crop_and_resize) -- TEST INFRASTRUCTURE.
 *
 * Follows third_party/RoIAlign.pytorch/roi_align/src/crop_and_resize.cpp:7-113 (CropAndResizePerBox): per box
 * (y1,x1,y2,x2 normalised by (H-1),(W-1)), sample crop_h x crop_w points, bilinear from floor/ceil neighbours,
 * extrapolation value outside [0, H-1] x [0, W-1].  Plain C, single thread, same float operation order.
 * Pinned against the reference's known-answer vector (README.md:42-96) in tests/test_oracle_cpu.py.
 */
#include <math.h>

int roialign_oracle_forward(const float* image, int batch, int depth, int ih, int iw, const float* boxes, const int* box_index,
                            int num_boxes, float extrapolation_value, int crop_h, int crop_w, float* crops) {
    const long image_channel_elements = (long)ih * iw;
    const long image_elements = depth * image_channel_elements;
    const long channel_elements = (long)crop_h * crop_w;
    const long crop_elements = depth * channel_elements;
    for (int b = 0; b < num_boxes; ++b) {
        const float* box = boxes + b * 4;
        const float y1 = box[0], x1 = box[1], y2 = box[2], x2 = box[3];
        const int b_in = box_index[b];
        if (b_in < 0 || b_in >= batch) return -1;
        const float height_scale = (crop_h > 1) ? (y2 - y1) * (ih - 1) / (crop_h - 1) : 0;
        const float width_scale = (crop_w > 1) ? (x2 - x1) * (iw - 1) / (crop_w - 1) : 0;
        for (int y = 0; y < crop_h; ++y) {
            const float in_y = (crop_h > 1) ? y1 * (ih - 1) + y * height_scale : 0.5 * (y1 + y2) * (ih - 1);
            if (in_y < 0 || in_y > ih - 1) {
                for (int x = 0; x < crop_w; ++x)
                    for (int d = 0; d < depth; ++d) crops[crop_elements * b + channel_elements * d + y * crop_w + x] = extrapolation_value;
                continue;
            }
            const int top = (int)floorf(in_y), bottom = (int)ceilf(in_y);
            const float y_lerp = in_y - top;
            for (int x = 0; x < crop_w; ++x) {
                const float in_x = (crop_w > 1) ? x1 * (iw - 1) + x * width_scale : 0.5 * (x1 + x2) * (iw - 1);
                if (in_x < 0 || in_x > iw - 1) {
                    for (int d = 0; d < depth; ++d) crops[crop_elements * b + channel_elements * d + y * crop_w + x] = extrapolation_value;
                    continue;
                }
                const int left = (int)floorf(in_x), right = (int)ceilf(in_x);
                const float x_lerp = in_x - left;
                for (int d = 0; d < depth; ++d) {
                    const float* p = image + b_in * image_elements + d * image_channel_elements;
                    const float tl = p[(long)top * iw + left], tr = p[(long)top * iw + right];
                    const float bl = p[(long)bottom * iw + left], br = p[(long)bottom * iw + right];
                    const float t = tl + (tr - tl) * x_lerp;
                    const float bt = bl + (br - bl) * x_lerp;
                    crops[crop_elements * b + channel_elements * d + y * crop_w + x] = t + (bt - t) * y_lerp;
                }
            }
        }
    }
    return 0;
}
