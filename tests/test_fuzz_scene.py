"""Short run of tools/fuzz_scene_files.py: corrupt scene.json / Scene.bin files are rejected or built, never crashed on."""
import importlib.util
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def test_corrupt_scene_files_are_rejected_or_built(akr, tmp_path):
    spec = importlib.util.spec_from_file_location("fuzz_scene_files", os.path.join(os.path.dirname(HERE), "tools", "fuzz_scene_files.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    built, build_rejected, load_rejected = fz.run(300, 5, str(tmp_path))
    assert built + build_rejected + load_rejected == 300
    assert built > 30 and build_rejected > 10 and load_rejected > 50
