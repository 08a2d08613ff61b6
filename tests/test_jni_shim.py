"""The JNI shim (integration/jni/lagassign_jni.c) built against a minimal stand-in for
<jni.h> and driven through a fake JNIEnv (no JDK exists in this image): error -> exception
mapping without a GPU, and the reference's README example through assignBatch on a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PKG = os.path.join(ROOT, "kafka-lag-based-assignor_b200")
JNI = os.path.join(ROOT, "integration", "jni")
EXE = os.path.join(JNI, "fake", "harness")


@pytest.fixture(scope="module")
def harness():
    import __graft_entry__
    __graft_entry__.build()
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    cmd = [cc, "-O1", "-Wall", "-I", os.path.join(JNI, "fake"), "-I", os.path.join(ROOT, "include"),
           "-o", EXE, os.path.join(JNI, "fake", "harness.c"), os.path.join(JNI, "lagassign_jni.c"),
           "-L", PKG, "-llagassign", f"-Wl,-rpath,{PKG}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def _has_gpu():
    import torch
    return torch.cuda.is_available()


@pytest.mark.skipif(_has_gpu(), reason="only meaningful without a GPU")
def test_jni_create_without_gpu_throws(harness):
    r = subprocess.run([harness, "nogpu"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "IllegalStateException" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_jni_assign_batch_on_gpu(harness):
    r = subprocess.run([harness, "gpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "README example through the JNI shim" in r.stdout and "RuntimeException" in r.stdout
