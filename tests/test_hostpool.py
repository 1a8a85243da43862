"""The worker pool of the packed record transfers (dl-poly_b200/csrc/hostpool.h) is plain C++: build a stress driver with the host
compiler -- under the thread sanitizer when the toolchain has it -- and run the three ways hostio.cu uses the pool: chunks consumed in
order by the calling thread while workers produce them (upload), workers gated chunk by chunk by the calling thread (download), and
the zero-worker pool that runs everything inline."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = r'''
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include "hostpool.h"
using dlp_hostpool::HostPool;
using dlp_hostpool::MAXCH;

int main(int argc, char** argv) {
  const int rounds = argc > 1 ? atoi(argv[1]) : 200;
  long long checks = 0;
  for (int nt : {0, 1, 2, 3, 7}) {
    HostPool pool(nt);
    std::vector<long long> data(MAXCH * 1000), out(MAXCH);
    std::atomic<int> avail[MAXCH];
    for (int r = 0; r < rounds; ++r) {
      const int nch = 1 + (r * 7 + nt) % MAXCH, len = 1 + (r * 13) % 1000;
      // upload pattern: workers fill chunk c, the caller consumes the chunks in order as they complete
      pool.start(nch, [&, len, r](int c) { for (int i = 0; i < len; ++i) data[(size_t)c * 1000 + i] = (long long)r * 1000003 + c * 1009 + i; });
      for (int c = 0; c < nch; ++c) {
        pool.wait_chunk(c);
        long long s = 0;
        for (int i = 0; i < len; ++i) s += data[(size_t)c * 1000 + i];
        const long long want = (long long)len * ((long long)r * 1000003 + c * 1009) + (long long)len * (len - 1) / 2;
        if (s != want) { std::printf("upload pattern: chunk %d of round %d wrong\n", c, r); return 1; }
        ++checks;
      }
      pool.wait_all();
      // download pattern: the caller releases chunk c, a worker that drew c waits for the release and then reads what the caller wrote
      for (int c = 0; c < nch; ++c) avail[c].store(0, std::memory_order_relaxed);
      if (nt == 0) for (int c = 0; c < nch; ++c) { data[(size_t)c * 1000] = r + c; avail[c].store(1, std::memory_order_release); }   // inline pool: release first
      pool.start(nch, [&, r](int c) {
        unsigned spins = 0;
        while (avail[c].load(std::memory_order_acquire) == 0) HostPool::pause(++spins);
        out[c] = data[(size_t)c * 1000] - (r + c);
      });
      if (nt != 0) for (int c = 0; c < nch; ++c) { data[(size_t)c * 1000] = r + c; avail[c].store(1, std::memory_order_release); }
      pool.wait_all();
      for (int c = 0; c < nch; ++c) { if (out[c] != 0) { std::printf("download pattern: chunk %d of round %d wrong\n", c, r); return 1; } ++checks; }
    }
  }
  std::printf("hostpool ok %lld\n", checks);
  return 0;
}
'''


def _build(tmp_path, sanitize):
    src = tmp_path / "hostpool_stress.cpp"
    src.write_text(DRIVER)
    exe = tmp_path / ("hostpool_stress_tsan" if sanitize else "hostpool_stress")
    cmd = [os.environ.get("CXX", "g++"), "-std=c++17", "-O1", "-g", "-pthread", "-I", os.path.join(ROOT, "dl-poly_b200", "csrc"), str(src), "-o", str(exe)]
    if sanitize:
        cmd[4:4] = ["-fsanitize=thread"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return (exe if r.returncode == 0 else None), r.stdout


def test_hostpool_patterns_of_the_packed_transfers(tmp_path):
    if shutil.which(os.environ.get("CXX", "g++")) is None:
        pytest.skip("no host C++ compiler")
    exe, log = _build(tmp_path, sanitize=False)
    assert exe is not None, log
    r = subprocess.run([str(exe), "300"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "hostpool ok" in r.stdout, r.stdout[-2000:]


def test_hostpool_under_the_thread_sanitizer(tmp_path):
    if shutil.which(os.environ.get("CXX", "g++")) is None:
        pytest.skip("no host C++ compiler")
    exe, log = _build(tmp_path, sanitize=True)
    if exe is None:
        pytest.skip("the toolchain has no thread sanitizer runtime: %s" % log[-200:])
    r = subprocess.run([str(exe), "60"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600,
                       env=dict(os.environ, TSAN_OPTIONS="halt_on_error=1"))
    if "FATAL: ThreadSanitizer" in r.stdout and "unexpected memory mapping" in r.stdout:
        pytest.skip("thread sanitizer cannot run in this container (address space layout)")
    assert r.returncode == 0 and "hostpool ok" in r.stdout and "WARNING: ThreadSanitizer" not in r.stdout, r.stdout[-3000:]
