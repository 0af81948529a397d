"""The C++ host layer (ann_force_b200/plugin): same classes and entry points as the reference plugin, compiled
against the OpenMM stand-in.  CPU: it type-checks.  GPU: the reference-style test program runs through a Context."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
PLUGIN = os.path.join(ROOT, "ann_force_b200", "plugin")


def test_plugin_layer_type_checks():
    subprocess.run(["make", "-C", CPP, "check"], check=True, capture_output=True)


def test_plugin_keeps_the_reference_api_surface():
    """Every public method of the reference ANN_Force (openmmapi/include/openmm/ANN_Force.h:57-112), the kernel
    contract (ANN_Kernels.h:23-53) and the plugin entry points (CudaANN_KernelFactory.cpp:43-65) exist by name."""
    header = open(os.path.join(PLUGIN, "openmmapi", "include", "openmm", "ANN_Force.h")).read()
    swig = open(os.path.join(PLUGIN, "python", "ANN.i")).read()
    methods = ["get_num_of_nodes", "set_num_of_nodes", "get_index_of_backbone_atoms", "set_index_of_backbone_atoms",
               "get_coeffients_of_connections", "set_coeffients_of_connections", "get_layer_types", "set_layer_types",
               "get_values_of_biased_nodes", "set_values_of_biased_nodes", "get_potential_center", "set_potential_center",
               "get_force_constant", "set_force_constant", "get_scaling_factor", "set_scaling_factor",
               "get_data_type_in_input_layer", "set_data_type_in_input_layer",
               "get_list_of_index_of_atoms_forming_dihedrals", "set_list_of_index_of_atoms_forming_dihedrals",
               "set_list_of_index_of_atoms_forming_dihedrals_from_index_of_backbone_atoms",
               "get_list_of_pair_index_for_distances", "set_list_of_pair_index_for_distances", "updateParametersInContext"]
    for m in methods:
        assert re.search(r"\b%s\s*\(" % m, header), m
        assert re.search(r"\b%s\s*\(" % m, swig), m
    from ann_force_b200 import ANN_Force
    for m in methods[:-1]:
        assert hasattr(ANN_Force, m), m
    kernels = open(os.path.join(PLUGIN, "openmmapi", "include", "openmm", "ANN_Kernels.h")).read()
    assert '"CalcANN_Force"' in kernels
    for m in ["initialize", "execute", "copyParametersToContext"]:
        assert re.search(r"virtual\s+\w+\s+%s\s*\(" % m, kernels), m
    factory = open(os.path.join(PLUGIN, "platforms", "cuda", "src", "CudaANN_KernelFactory.cpp")).read()
    for sym in ["registerPlatforms", "registerKernelFactories", "registerANN_CudaKernelFactories"]:
        assert re.search(r'extern "C" OPENMM_EXPORT void %s\(\)' % sym, factory), sym


@pytest.mark.gpu
def test_plugin_layer_runs_the_reference_style_tests():
    subprocess.run(["make", "-C", CPP], check=True, capture_output=True)
    out = subprocess.run([os.path.join(CPP, "_build", "test_ann_plugin")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("Done"), out.stdout


def test_serialization_proxy_round_trip():
    """ANN_Force <-> SerializationNode, bit-exact, registered at library load (CPU only)."""
    exe = os.path.join(CPP, "_build", "test_serialization")
    subprocess.run(["make", "-C", CPP, exe], check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "serialization ok" in out.stdout, out.stdout + out.stderr


def test_cmake_packaging_builds_the_plugin_libraries(tmp_path):
    """The CMake project of the plugin (same targets and install layout as the reference's: libANN in lib/, libANN_CUDA in
    lib/plugins) configures, builds against the OpenMM stand-in, passes its ctest, installs, and the CUDA plugin exports
    the entry points OpenMM's plugin loader and static users call (CudaANN_KernelFactory.cpp:43-65 of the reference)."""
    import shutil
    if shutil.which("cmake") is None or not os.path.exists(os.path.join(ROOT, "ann_force_b200", "libannforce_b200.so")):
        pytest.skip("needs cmake and the built kernel library")
    build, prefix = str(tmp_path / "build"), str(tmp_path / "prefix")
    run = lambda *cmd, **kw: subprocess.run(cmd, check=True, capture_output=True, text=True, **kw)
    run("cmake", "-S", PLUGIN, "-B", build, "-DANN_USE_OPENMM_STUB=ON", "-DCMAKE_INSTALL_PREFIX=" + prefix)
    run("cmake", "--build", build, "-j4")
    assert "100% tests passed" in run("ctest", cwd=build).stdout
    run("cmake", "--install", build)
    assert os.path.exists(os.path.join(prefix, "lib", "libANN.so"))
    assert os.path.exists(os.path.join(prefix, "lib", "plugins", "libANN_CUDA.so"))
    assert os.path.exists(os.path.join(prefix, "include", "openmm", "ANN_Force.h"))
    symbols = run("nm", "-D", "--defined-only", os.path.join(prefix, "lib", "plugins", "libANN_CUDA.so")).stdout
    for name in ("registerPlatforms", "registerKernelFactories", "registerANN_CudaKernelFactories"):
        assert re.search(r"\bT %s\b" % name, symbols), name
    api = run("nm", "-D", "--defined-only", os.path.join(prefix, "lib", "libANN.so")).stdout
    assert re.search(r"\bT registerANNSerializationProxies\b", api)
