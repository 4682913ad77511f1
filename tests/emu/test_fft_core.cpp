#define DPX_EMU
#include "dpx_fft_core.cuh"
#include <complex>
#include <cstdio>
#include <random>
using namespace dpx::fft;
template <class T> double test() {
  std::mt19937 rng(1); std::uniform_real_distribution<float> U(-1, 1);
  constexpr int N = T::N, C = T::COLS;
  std::vector<std::complex<double>> x(N * C), X(N * C);
  for (auto& v : x) v = {U(rng), U(rng)};
  for (int c = 0; c < C; ++c) for (int k = 0; k < N; ++k) { std::complex<double> s = 0; for (int n = 0; n < N; ++n) s += x[n * C + c] * std::polar(1.0, -2 * M_PI * (double)((long)k * n % N) / N); X[k * C + c] = s; }
  std::vector<float2> tw(TwiddleLayout<T>::TOTAL);
  for (int j = 0; j < T::MA; ++j) for (int q = 0; q < T::RA; ++q) tw[TwiddleLayout<T>::A_OFF + j * T::RA + q] = make_float2((float)cos(2 * M_PI * j * q / N), (float)-sin(2 * M_PI * j * q / N));
  for (int j = 0; j < T::MB; ++j) for (int q = 0; q < T::RB; ++q) tw[TwiddleLayout<T>::B_OFF + j * T::RB + q] = make_float2((float)cos(2 * M_PI * j * q / T::MA), (float)-sin(2 * M_PI * j * q / T::MA));
  double err_f = 0, err_i = 0, nrm = 0;
  emu::launch(dim3(1), dim3(64), T::SMEM_FLOAT2 * sizeof(float2), [&]() {
    DPX_DYN_SMEM(float2, sm);
    int tid = threadIdx.x;
    for (int i = tid; i < N * C; i += 64) sm[T::phys(i / C, i % C)] = make_float2((float)x[i].real(), (float)x[i].imag());
    __syncthreads();
    tile_fft_forward<T>(sm, tw.data(), tid, 64);
    if (tid == 0) for (int k = 0; k < N; ++k) for (int c = 0; c < C; ++c) { float2 v = sm[T::phys(T::pos_of_freq(k), c)]; err_f += std::norm(std::complex<double>(v.x, v.y) - X[k * C + c]); nrm += std::norm(X[k * C + c]);
        if (T::freq_of_pos(T::pos_of_freq(k)) != k) printf("pos/freq mismatch\n"); }
    __syncthreads();
    tile_fft_inverse<T>(sm, tw.data(), tid, 64);
    if (tid == 0) for (int i = 0; i < N * C; ++i) { float2 v = sm[T::phys(i / C, i % C)]; err_i += std::norm(std::complex<double>(v.x, v.y) / (double)N - x[i]); }
  });
  printf("N=%d (%d,%d,%d) cols=%d  fwd rel err %.3e  roundtrip err %.3e\n", N, T::RA, T::RB, T::RC, C, sqrt(err_f / nrm), sqrt(err_i / (N * C)));
  return sqrt(err_f / nrm);
}
int main() {
  test<Tile<256, 8, 8, 4, 2>>(); test<Tile<512, 8, 8, 8, 4>>(); test<Tile<1024, 16, 8, 8, 2>>(); test<Tile<2048, 16, 16, 8, 4>>(); test<Tile<4096, 16, 16, 16, 2>>();
  test<Tile<64, 4, 4, 4, 4>>(); test<Tile<128, 8, 4, 4, 2>>();
}
