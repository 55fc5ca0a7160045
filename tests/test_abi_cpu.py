"""CPU: the C-ABI library loads, exports every symbol include/vslam_b200.h declares, and refuses to run without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from visualslam_android_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "vslam_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vslam_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = api.load()
    names = _header_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/vslam_b200.h but not exported"
    assert sorted(api.ABI_SYMBOLS) == names, "api.ABI_SYMBOLS must list exactly the header's entry points"


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(api.VslamError) as e:
        api.Context(640, 480)
    assert e.value.code == api.E_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_create_rejects_bad_geometry():
    L = api.load()
    cfg = api.Config(); L.vslam_default_config(C.byref(cfg))
    h = C.c_void_p()
    for w, hh, p in ((650, 480, 11), (640, 481, 11), (640, 480, 12), (0, 0, 11)):
        cfg.width, cfg.height, cfg.patch_size = w, hh, p
        assert L.vslam_create(C.byref(cfg), C.byref(h)) == api.E_INVALID
        assert L.vslam_last_error(None)


def test_default_params_are_the_reference_constants():
    L = api.load()
    p = api.Params(); L.vslam_default_params(C.byref(p))
    # jni/Tracker.cc:405-410, 495-497, 518
    assert (p.coarse_min, p.coarse_max, p.coarse_range, p.coarse_subpix_its, p.coarse_min_vel) == (20, 60, 30, 8, 0.006)
    assert (p.fine_range, p.fine_range_after_coarse, p.fine_subpix_its_top_level, p.max_patches_per_frame, p.use_sbi) == (10, 5, 8, 1000, 1)
    # execution-only knobs: one stream group, parallel normal-equation sums, the round-2 kernels, frame look-ahead left to the library
    assert (p.stream_groups, p.serial_normal_equations, p.pose_kernel, p.search_kernel, p.frame_lookahead, p.coarse_chain) == (1, 0, 0, 0, -1, -1)


def test_params_struct_matches_the_header(tmp_path):
    """The ctypes mirror of vslam_params has the C struct's size (a field added to the header but not to api.Params would let
    vslam_default_params write past the Python object)."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include "vslam_b200.h"\n#include <stdio.h>\nint main(void) { printf("%zu %zu\\n", sizeof(vslam_params), sizeof(vslam_config)); return 0; }\n')
    exe = tmp_path / "sz"
    r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    sp, sc = map(int, subprocess.run([str(exe)], capture_output=True, text=True).stdout.split())
    assert sp == C.sizeof(api.Params) and sc == C.sizeof(api.Config)


@pytest.mark.parametrize("size", [(640, 480), (1920, 1080), (3840, 2160)])
def test_camera_from_params_matches_refresh_params(size):
    w, h = size
    fixed = api.camera_from_params(synth.CAMERA_PARAMS, w, h, as_shipped_radius=False)
    assert np.array_equal(fixed, synth.Camera(w, h).scalars())
    shipped = api.camera_from_params(synth.CAMERA_PARAMS, w, h, as_shipped_radius=True)
    assert shipped[9] == 0.0 and shipped[10] == 0.0     # as shipped every projection is rejected
    assert np.array_equal(shipped[:9], fixed[:9])


def test_cpp_shell_compiles_against_the_reference_type_surface(tmp_path):
    """include/vslam_b200_shell.hpp (KeyFrame / Tracker API over the C-ABI) compiles and links with the stand-in cv::Mat / Eigen headers."""
    import subprocess
    src = tmp_path / "shell_check.cc"
    src.write_text('#include "vslam_b200_shell.hpp"\n'
                   'int main() { double cam[13]; double p5[5] = {0.841906, 1.10893, 0.505171, 0.470265, -0.0133843};\n'
                   '  vslam_camera_from_params(p5, 640, 480, 0, cam);\n'
                   '  try { vslam_b200::Context ctx(640, 480, 1, 16); vslam_b200::Tracker t(ctx, 0, cam); vslam_b200::KeyFrame kf(ctx, 0);\n'
                   '        cv::Mat g(480, 640, CV_8UC1), c(1, 1, CV_8UC4); kf.MakeKeyFrame_Lite(g, c); t.TrackFrame(g, c, false); (void)t.GetCurrentPose(); }\n'
                   '  catch (const std::exception& e) { return 3; }   // no GPU here: vslam_create fails loudly\n'
                   '  return 0; }\n')
    exe = tmp_path / "shell_check"
    libdir = os.path.join(ROOT, "visualslam_android_b200")
    cmd = ["g++", "-std=gnu++11", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "shim"), str(src), "-o", str(exe),
           "-L", libdir, "-lvslam_b200", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    import torch
    rc = subprocess.run([str(exe)]).returncode
    assert rc == (0 if torch.cuda.is_available() else 3)


def test_header_is_plain_c(tmp_path):
    """include/vslam_b200.h is a C header (the boundary a cgo / JNI / ctypes binding sees): it compiles as C99 and a C translation unit
    that drives the multi-camera loop of INTEGRATION.md links against the library."""
    import subprocess
    src = tmp_path / "abi_check.c"
    src.write_text('#include "vslam_b200.h"\n'
                   '#include <stdio.h>\n'
                   'int main(void) {\n'
                   '  vslam_config cfg; vslam_params prm; vslam_ctx* ctx = 0; double cam[13], p5[5] = {0.841906, 1.10893, 0.505171, 0.470265, -0.0133843};\n'
                   '  vslam_default_config(&cfg); vslam_default_params(&prm);\n'
                   '  cfg.n_streams = 4; cfg.max_points = 100;\n'
                   '  vslam_camera_from_params(p5, cfg.width, cfg.height, 0, cam);\n'
                   '  if (vslam_create(&cfg, &ctx) != VSLAM_OK) { printf("%s\\n", vslam_last_error(0)); return 3; }\n'
                   '  vslam_set_camera(ctx, cam); vslam_set_params(ctx, &prm);\n'
                   '  vslam_destroy(ctx);\n'
                   '  return 0; }\n')
    exe = tmp_path / "abi_check"
    libdir = os.path.join(ROOT, "visualslam_android_b200")
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", libdir, "-lvslam_b200",
                        f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    import torch
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == (0 if torch.cuda.is_available() else 3), run.stdout
    if run.returncode == 3:
        assert "no CUDA device" in run.stdout


def test_map_file_header_is_readable_without_a_gpu(tmp_path):
    """vslam_map_file_info is host-only: it parses a file written by the tests' independent restatement of the layout
    (tests/common.py mapfile_pack) and rejects files with a wrong magic / version or a short header."""
    import common
    from visualslam_android_b200 import api
    rs = np.random.RandomState(3)
    W, H, n = 64, 32, 5
    cam13 = rs.rand(13)
    pts = dict(world=rs.randn(n, 3), right=rs.randn(n, 3), down=rs.randn(n, 3), ircenter=rs.randint(0, 30, (n, 2)), srclevel=rs.randint(0, 4, n), srckf=np.zeros(n, dtype=np.int32))
    kfs = [(0, 0, rs.randint(0, 255, (H, W)).astype(np.uint8)), (2, -1, rs.randint(0, 255, (H, W)).astype(np.uint8))]
    blob = common.mapfile_pack(W, H, cam13, kfs, pts, (np.array([0], dtype=np.int32), rs.randn(1, 12)))
    back = common.mapfile_unpack(blob)
    assert np.array_equal(back["points"]["world"], pts["world"]) and back["keyframes"][1][0] == 2 and np.array_equal(back["keyframes"][1][2], kfs[1][2])
    path = tmp_path / "m.vsmap"
    path.write_bytes(blob)
    info = api.map_file_info(path)
    assert (info["width"], info["height"], info["n_points"], info["n_keyframes"], info["n_reloc_keyframes"]) == (W, H, n, 2, 1)
    assert np.array_equal(info["cam13"], cam13)
    for bad in (b"NOTAMAP!" + blob[8:], blob[:8] + b"\x02\x00\x00\x00" + blob[12:], blob[:100]):
        path.write_bytes(bad)
        with pytest.raises(api.VslamError) as e:
            api.map_file_info(path)
        assert e.value.code == api.E_INVALID
    with pytest.raises(api.VslamError) as e:
        api.map_file_info(tmp_path / "does_not_exist")
    assert e.value.code == api.E_IO
