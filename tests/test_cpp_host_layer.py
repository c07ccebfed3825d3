"""The header-only C++ host layer (include/mppi_b200/): builds with plain g++ against the Eigen shim, links to the C-ABI
library, and runs the reference's cartpole example flow (tests/cpp/cartpole_example.cpp). On a box without a GPU the
binary must stop at the C-ABI's NO_DEVICE error (exit code 5) — never a CPU fallback; on the B200 it must swing up."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "cartpole_example.bin")


def _build():
    lib_dir = os.path.join(ROOT, "mppi-generic_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wno-unused-variable", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "cartpole_example.cpp"), "-o", EXE, "-L", lib_dir,
                           "-lmppi_b200", "-Wl,-rpath," + lib_dir])


def test_host_layer_compiles_and_fails_loudly_without_a_device():
    _build()
    p = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    if p.returncode == 5:
        assert "no CUDA device" in p.stdout
    else:  # a GPU is present: the example must succeed
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]


@pytest.mark.gpu
def test_host_layer_cartpole_example_swings_up_on_the_gpu():
    _build()
    p = subprocess.run([EXE], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "final pole angle error" in p.stdout


RACER_EXE = os.path.join(ROOT, "tests", "cpp", "racer_colored_example.bin")


def _build_racer():
    lib_dir = os.path.join(ROOT, "mppi-generic_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wno-unused-variable", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "racer_colored_example.cpp"), "-o", RACER_EXE, "-L", lib_dir,
                           "-lmppi_b200", "-Wl,-rpath," + lib_dir])


def test_racer_colored_example_compiles_against_reference_include_paths():
    """C5 through the C++ layer via the reference's own include paths (include/mppi/**.cuh forwarders)."""
    _build_racer()
    p = subprocess.run([RACER_EXE], capture_output=True, text=True, timeout=600)
    if p.returncode == 5:
        assert "no CUDA device" in p.stdout
    else:
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]


@pytest.mark.gpu
def test_racer_colored_example_tracks_speed_on_the_gpu():
    _build_racer()
    p = subprocess.run([RACER_EXE], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "speed after 80 steps" in p.stdout


QUAD_EXE = os.path.join(ROOT, "tests", "cpp", "quadrotor_example.bin")


def _build_quad():
    lib_dir = os.path.join(ROOT, "mppi-generic_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wno-unused-variable", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "quadrotor_example.cpp"), "-o", QUAD_EXE, "-L", lib_dir,
                           "-lmppi_b200", "-Wl,-rpath," + lib_dir])


def test_quadrotor_example_compiles_against_reference_include_paths():
    """Quadrotor pair through the C++ layer via instantiations/quadrotor_mppi/quadrotor_mppi.cuh."""
    _build_quad()
    p = subprocess.run([QUAD_EXE], capture_output=True, text=True, timeout=600)
    if p.returncode == 5:
        assert "no CUDA device" in p.stdout
    else:
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]


@pytest.mark.gpu
def test_quadrotor_example_flies_to_the_goal_on_the_gpu():
    _build_quad()
    p = subprocess.run([QUAD_EXE], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "distance to goal after 250 steps" in p.stdout


# ---- the reference's instantiation libraries (src/controllers/*/: SURVEY §8 f4) ------------------------------------------
INST_EXE = os.path.join(ROOT, "tests", "cpp", "instantiation_user.bin")


def _build_instantiation_user():
    lib_dir = os.path.join(ROOT, "mppi-generic_b200")
    subprocess.check_call(["bash", os.path.join(ROOT, "src", "controllers", "build.sh")])
    obj = INST_EXE + ".o"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-DMPPIB_USE_INSTANTIATION_LIBRARY", "-I", os.path.join(ROOT, "include"),
                           "-c", os.path.join(ROOT, "tests", "cpp", "instantiation_user.cpp"), "-o", obj])
    subprocess.check_call(["g++", obj, "-o", INST_EXE, "-L", lib_dir, "-l:libmppi_b200_controllers.so", "-l:libmppi_b200.so",
                           "-Wl,-rpath," + lib_dir])
    return obj


def test_instantiation_library_provides_the_prebuilt_controllers():
    """A translation unit that includes <mppi/instantiations/cartpole_mppi/cartpole_mppi.cuh> with
    MPPIB_USE_INSTANTIATION_LIBRARY does not instantiate the controller: computeControl is an undefined symbol of the object
    and a defined one of libmppi_b200_controllers.so, for every class the reference's src/controllers/*/ pre-build."""
    obj = _build_instantiation_user()
    und = subprocess.run(["nm", "-C", "--undefined-only", obj], capture_output=True, text=True, check=True).stdout
    assert "VanillaMPPIController<CartpoleDynamics, CartpoleQuadraticCost, DDPFeedback<CartpoleDynamics, 100>, 100, 2048" in und
    assert "::computeControl(" in und
    lib = os.path.join(ROOT, "mppi-generic_b200", "libmppi_b200_controllers.so")
    defined = subprocess.run(["nm", "-DC", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    for needle in ("VanillaMPPIController<CartpoleDynamics, CartpoleQuadraticCost, DDPFeedback<CartpoleDynamics, 100>, 100, 2048",
                   "VanillaMPPIController<CartpoleDynamics, CartpoleQuadraticCost, DDPFeedback<CartpoleDynamics, 100>, 100, 256",
                   "VanillaMPPIController<CartpoleDynamics, CartpoleQuadraticCost, DDPFeedback<CartpoleDynamics, 150>, 150, 512",
                   "VanillaMPPIController<DoubleIntegratorDynamics, DoubleIntegratorCircleCost, DDPFeedback<DoubleIntegratorDynamics, 50>, 50, 1024",
                   "TubeMPPIController<DoubleIntegratorDynamics, DoubleIntegratorCircleCost, DDPFeedback<DoubleIntegratorDynamics, 100>, 100, 1024",
                   "VanillaMPPIController<QuadrotorDynamics, QuadrotorQuadraticCost, DDPFeedback<QuadrotorDynamics, 100>, 100, 512",
                   "VanillaMPPIController<NeuralNetModel<7, 2, 3>, ARStandardCost, DDPFeedback<NeuralNetModel<7, 2, 3>, 150>, 150, 1920"):
        assert any(needle in ln and "::computeControl(" in ln for ln in defined.splitlines()), needle
    p = subprocess.run([INST_EXE], capture_output=True, text=True, timeout=600)
    if p.returncode != 0:  # no GPU here: the host layer reports it like the reference's HANDLE_ERROR (message + exit)
        assert "no CUDA device" in (p.stdout + p.stderr), p.stdout[-1000:] + p.stderr[-1000:]


@pytest.mark.gpu
def test_instantiation_library_user_runs_on_the_gpu():
    _build_instantiation_user()
    p = subprocess.run([INST_EXE], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]


def test_texture_helper_through_the_cpp_layer():
    """TwoDTextureHelper<float> (include/mppi/utils/texture_helpers/two_d_texture_helper.cuh forwarder): the reference's own
    world-pose known answers and a host step of the RACER model over a sloped elevation map (tests/cpp/texture_helper_test.cpp)."""
    lib_dir = os.path.join(ROOT, "mppi-generic_b200")
    exe = os.path.join(ROOT, "tests", "cpp", "texture_helper_test.bin")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wno-unused-variable", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "texture_helper_test.cpp"), "-o", exe, "-L", lib_dir,
                           "-l:libmppi_b200.so", "-Wl,-rpath," + lib_dir])
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
