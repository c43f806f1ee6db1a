// extern "C" shim around the REFERENCE's own correlation launchers (test infrastructure only).
// The launchers are compiled, unmodified, from /root/reference/Nets/Native/shift_corr.cu.cc by build_ref.sh
// (they have C++ linkage there: shift_corr.cu.cc:193-194).  Only the forward launcher is wrapped: the
// reference backward is defective and is never used as an oracle.
void ShiftCorrKernelLauncher(const float* values0, const float* values1, const int max_disp, const int batch_size,
                             const int in_h, const int in_w, const int in_channels, float* out);

extern "C" __attribute__((visibility("default"))) int ref_shift_corr(const float* in0_padded, const float* in1_padded,
                                                                      int max_disp, int batch, int h, int w_padded,
                                                                      int channels, float* out_nchw) {
    // legacy default stream, exactly as the reference launches it (shift_corr.cu.cc:226)
    ShiftCorrKernelLauncher(in0_padded, in1_padded, max_disp, batch, h, w_padded, channels, out_nchw);
    return (int)cudaGetLastError();
}
