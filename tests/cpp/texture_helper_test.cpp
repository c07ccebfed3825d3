// TwoDTextureHelper<float> through the C++ host layer via the reference's include paths: the reference's own known answers
// (tests/texture_helpers/two_d_texture_helper_test.cu:451-541) and a host step of the RACER model over a sloped map.
#include <mppi/dynamics/racer_dubins/racer_dubins_elevation_lstm_steering.cuh>
#include <mppi/utils/texture_helpers/two_d_texture_helper.cuh>
#include <cstdio>
int main()
{
  TwoDTextureHelper<float> helper(1);
  cudaExtent extent = make_cudaExtent(10, 20, 0);
  helper.setExtent(0, extent);
  std::vector<float> data(200);
  for (int i = 0; i < 200; i++) data[i] = i;
  std::array<float3, 3> rot{};
  rot[0] = make_float3(0, 1, 0); rot[1] = make_float3(1, 0, 0); rot[2] = make_float3(0, 0, 1);
  helper.updateRotation(0, rot);
  helper.updateOrigin(0, make_float3(1, 2, 3));
  helper.updateTexture(0, data);
  helper.updateResolution(0, 10);
  helper.enableTexture(0);
  helper.copyToDevice(true);
  // two_d_texture_helper_test.cu:451-541 — query (0.5, 0) -> 4.5 ; (0, 0.5) -> 95
  float a = helper.queryTextureAtWorldPose(0, make_float3(0.0f * 200 + 1, 0.5f * 100 + 2, 3));
  float b = helper.queryTextureAtWorldPose(0, make_float3(0.5f * 200 + 1, 0.0f * 100 + 2, 3));
  printf("%f %f\n", a, b);
  if (!(a == 4.5f && b == 95.0f))
    return 1;
  // the model's own helper: a plane rising 0.1 m per metre along x -> the vehicle pitches nose-up (negative pitch: rear lower)
  std::vector<int> init_layers = { 23, 100, 8 }, out_layers = { 8, 20, 1 };  // the reference test architecture (:26-32)
  RacerDubinsElevationLSTMSteering model(3, 20, init_layers, 4, 4, out_layers, 11);
  TwoDTextureHelper<float>* tex = model.getTextureHelper();
  cudaExtent e2 = make_cudaExtent(40, 40, 0);
  std::vector<float> plane(1600);
  for (int i = 0; i < 40; i++)
    for (int j = 0; j < 40; j++)
      plane[i * 40 + j] = 0.1f * ((j + 0.5f) * 0.5f - 5.0f);
  tex->updateTexture(0, plane, e2);
  tex->updateResolution(0, 0.5f);
  tex->updateOrigin(0, make_float3(-5, -10, 0));
  tex->enableTexture(0);
  tex->copyToDevice();
  RacerDubinsElevationLSTMSteering::state_array x = RacerDubinsElevationLSTMSteering::state_array::Zero(), xn, xd;
  RacerDubinsElevationLSTMSteering::control_array u = RacerDubinsElevationLSTMSteering::control_array::Zero();
  RacerDubinsElevationLSTMSteering::output_array y;
  x(0) = 2.0f;
  model.step(x, xn, xd, u, y, 0.0f, 0.02f);
  const float expect_pitch = asinf(-0.1f * 2.981f / 2.981f);
  printf("pitch %f expect %f height %f\n", xn(7), expect_pitch, y(4));
  if (!(fabsf(xn(7) - expect_pitch) < 1e-4f && fabsf(xn(6)) < 1e-5f))
    return 2;
  // the init network: the reference's known answer (tests/nn_helpers/lstm_lstm_helper_test.cu:161-180): all weights 1, a
  // buffer of ones -> hidden = cell = 101
  std::vector<int> il = { 68, 100, 20 }, ol = { 14, 20, 1 };
  RacerDubinsElevationLSTMSteering m2(8, 60, il, 4, 10, ol, 6);
  m2.setAllValuesInit(std::vector<float>(4 * 60 * 60 + 4 * 60 * 8 + 6 * 60, 1.0f),
                      std::vector<float>(68 * 100 + 100 + 100 * 20 + 20, 1.0f));
  std::vector<float> ones(8 * 10, 1.0f);
  m2.initializeLSTM(ones.data(), 8, 10);
  const std::vector<float>& th = m2.getTheta();
  const size_t base = 4 * 10 * 10 + 4 * 10 * 4 + 4 * 10;
  printf("init hidden %f cell %f\n", th[base], th[base + 10]);
  return (th[base] == 101.0f && th[base + 19] == 101.0f) ? 0 : 3;
}
