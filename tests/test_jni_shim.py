"""The JNI shim (gigapaxos_b200/jni/gpx_jni.c) is the binding a gigapaxos maintainer compiles next to the jar
(INTEGRATION.md).  No JDK exists in this image (profiles/r2_java_probe_gpu_box.txt), so the shim cannot run here; what
can be checked is that it compiles against include/gpx.h with a stand-in jni.h declaring only the JNI calls it uses
(so a signature drift between the header and the shim is a compile error), that it links against libgpx.so with no
unresolved gpx_* symbol, and that it binds the entry points INTEGRATION.md names."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "gigapaxos_b200", "jni", "gpx_jni.c")

JNI_STUB = r"""
#ifndef STUB_JNI_H
#define STUB_JNI_H
#include <stdint.h>
typedef int32_t jint; typedef int64_t jlong; typedef double jdouble; typedef int32_t jsize;
typedef void* jobject; typedef jobject jclass; typedef jobject jstring; typedef jobject jlongArray;
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
  void* (*GetDirectBufferAddress)(JNIEnv*, jobject);
  jlong (*GetDirectBufferCapacity)(JNIEnv*, jobject);
  const char* (*GetStringUTFChars)(JNIEnv*, jstring, unsigned char*);
  void (*ReleaseStringUTFChars)(JNIEnv*, jstring, const char*);
  jstring (*NewStringUTF)(JNIEnv*, const char*);
  void (*SetLongArrayRegion)(JNIEnv*, jlongArray, jsize, jsize, const jlong*);
};
#endif
"""

EXPECTED = ["create", "destroy", "lastError", "createGroups", "destroyGroups", "dumpRows", "loadRows", "patch",
            "roundSubmit", "roundWait", "propose", "handleAccepts", "handleAcceptReplies", "handleDecisions",
            "handlePrepares", "handlePrepareReplies", "pauseGroups", "selectGroups", "clearGroupFlags", "missingDecisions", "logDrainAsync", "logDrainWait", "logRelease", "logRead", "logFind", "logGather", "getCpi", "getCounters",
            "spreadUniqueId", "spreadPlanNode", "spreadCreate", "spreadRound", "spreadDropped", "spreadDestroy"]


@pytest.fixture(scope="module")
def shim_obj(tmp_path_factory):
    d = tmp_path_factory.mktemp("jni")
    with open(d / "jni.h", "w") as f:
        f.write(JNI_STUB)
    obj = str(d / "gpx_jni.o")
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-Wno-unused-parameter", "-fPIC", "-c", "-I", str(d),
                        "-I", os.path.join(ROOT, "include"), SHIM, "-o", obj], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return obj


def test_shim_compiles_and_exports_the_binding(shim_obj):
    syms = subprocess.run(["nm", "-g", "--defined-only", shim_obj], capture_output=True, text=True).stdout
    have = set(re.findall(r"Java_edu_umass_cs_gigapaxos_gpx_PaxosEngine_(\w+)", syms))
    assert have == set(EXPECTED), (sorted(have - set(EXPECTED)), sorted(set(EXPECTED) - have))


def test_shim_calls_only_declared_entry_points(shim_obj):
    """every gpx_* the shim references is declared in include/gpx.h and exported by libgpx.so"""
    und = subprocess.run(["nm", "-u", shim_obj], capture_output=True, text=True).stdout
    used = set(re.findall(r"\b(gpx_\w+)", und))
    assert used, "the shim calls into libgpx"
    hdr = open(os.path.join(ROOT, "include", "gpx.h")).read()
    for s in used:
        assert re.search(r"\b%s\(" % s, hdr), s
    so = os.path.join(ROOT, "gigapaxos_b200", "libgpx.so")
    if os.path.exists(so):
        exp = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
        exported = set(re.findall(r"\b(gpx_\w+)", exp))
        assert used <= exported, sorted(used - exported)


JAVA_TO_JNI = {"long": "jlong", "int": "jint", "double": "jdouble", "ByteBuffer": "jobject", "String": "jstring",
               "long[]": "jlongArray", "void": "void"}


def test_java_class_and_shim_agree_on_every_native_method():
    java = open(os.path.join(ROOT, "gigapaxos_b200", "jni", "PaxosEngine.java")).read()
    c = open(SHIM).read()
    jm = {}
    for ret, name, args in re.findall(r"static native ([\w\[\]]+) (\w+)\(([^)]*)\);", java):
        types = [" ".join(a.split()[:-1]) for a in args.split(",")] if args.strip() else []
        jm[name] = (JAVA_TO_JNI[ret], [JAVA_TO_JNI[t] for t in types])
    cm = {}
    for ret, name, args in re.findall(r"JNIEXPORT (\w+) JNICALL GPX_JNI\((\w+)\)\(JNIEnv\* env, jclass cls([^)]*)\)", c):
        types = [a.split()[0] for a in args.split(",") if a.strip()]
        cm[name] = (ret, types)
    assert set(jm) == set(cm) == set(EXPECTED)
    for name in EXPECTED:
        assert jm[name] == cm[name], (name, jm[name], cm[name])
    assert "package edu.umass.cs.gigapaxos.gpx;" in java and "class PaxosEngine" in java  # = the GPX_JNI() prefix


def test_integration_doc_names_the_shim_entry_points():
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for name in ("roundSubmit", "roundWait", "createGroups", "logDrainAsync", "spreadRound", "handleAccepts"):
        assert name in doc, name
