// Host-layer check of the npz inputs: NeuralNetModel::loadParams(path) and ARStandardCost::loadTrackData(path) written
// against the reference's include paths. Prints key=value pairs that tests/test_npz_io.py compares with numpy.
#include <mppi/cost_functions/autorally/ar_standard_cost.cuh>
#include <mppi/dynamics/autorally/ar_nn_model.cuh>
#include <mppi/dynamics/racer_dubins/racer_dubins_elevation_lstm_steering.cuh>

#include <cmath>
#include <cstdio>

int main(int argc, char** argv)
{
  if (argc < 3)
    return 64;
  NeuralNetModel<7, 2, 3> model;
  model.loadParams(argv[1]);
  double sum = 0, abs_sum = 0;
  for (float v : model.getTheta())
    sum += v, abs_sum += std::fabs(v);
  printf("theta_sum=%.9g theta_abs=%.9g\n", sum, abs_sum);
  ARStandardCost cost;
  std::vector<float4> tex = cost.loadTrackData(argv[2]);
  double c0 = 0, c2 = 0;
  for (const float4& t : tex)
    c0 += t.x, c2 += t.z;
  printf("width=%d height=%d ch0_sum=%.9g ch2_sum=%.9g\n", cost.getWidth(), cost.getHeight(), c0, c2);
  int rc = 0;
  try
  {
    NeuralNetModel<7, 2, 3> other;
    other.loadParams("/nonexistent/model.npz");
  }
  catch (const std::exception& e)
  {
    rc = 1;
  }
  printf("missing_rc=%d\n", rc);
  if (argc >= 4)
  {  // prediction LSTM of the RACER model (H = 4, head {8, 20, 1}) from a PyTorch-layout npz
    std::vector<int> init_layers = { 23, 100, 8 }, out_layers = { 8, 20, 1 };
    RacerDubinsElevationLSTMSteering racer(3, 20, init_layers, 4, 4, out_layers, 11);
    racer.loadParamsLSTM(argv[3]);
    double ls = 0, la = 0;
    for (float v : racer.getTheta())
      ls += v, la += std::fabs(v);
    printf("lstm_sum=%.9g lstm_abs=%.9g\n", ls, la);
  }
  return 0;
}
