"""Host-side logic that needs no GPU: the C ABI surface, the loaders / plan compiler inside the
library (rs_model_check / rs_graph_check), the Python mirror of the reference's transcriber API."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from rhasspy_speech_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(lib):
    """Every function include/rs_b200.h declares is exported by librs_b200.so and bound."""
    with open(os.path.join(ROOT, "include", "rs_b200.h")) as f:
        hdr = f.read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rs_[a-z_0-9]+)\s*\(", hdr))
    bound = {name for name, _, _ in lib.SYMBOLS}
    assert declared == bound, declared ^ bound
    dll = ctypes.CDLL(lib.LIB_PATH)
    for name in declared:
        assert hasattr(dll, name), name


def test_no_cpu_fallback(lib, tiny_model):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lib.RsError, match="no CUDA device"):
        lib.Model(tiny_model.final_mdl, tiny_model.online_conf, 0)
    with pytest.raises(lib.RsError, match="no CUDA device"):
        lib.Graph(tiny_model.hclg, tiny_model.words_txt, 0)


def test_model_loader_and_plan(lib, tiny_model, synth):
    txt = lib.model_check(tiny_model.final_mdl, tiny_model.online_conf)
    head = txt.splitlines()[0]
    assert "pdfs %d " % tiny_model.num_pdfs in head and " sf 3 " in head and "ivector_dim 30" in head
    # context of the TINY net: lda (+-1) and tdnnf strides 1, 0, 3, 3  => 1 + 7 = 8 on each side
    assert "context -8/+8" in txt and "align 3" in txt
    # every TDNN-F layer is one GEMM launch with the ReLU, BatchNorm and bypass fused in
    assert txt.count("gemm tdnnf") == 8 and txt.count("addscaled(0.66") == 4
    assert "elementwise" not in txt and "logsoftmax" not in txt      # the xent branch is not computed
    assert "uttgemm lda.utt" in txt                                  # iVector part of the first layer


def test_model_loader_text_format(lib, synth, tmp_path):
    """Kaldi text-mode files parse to the same plan as binary ones."""
    import dataclasses
    spec = dataclasses.replace(synth.TINY, binary=False, priors=True)
    p = synth.write_model(str(tmp_path), spec)
    txt = lib.model_check(p.final_mdl, p.online_conf)
    assert "priors %d" % p.num_pdfs in txt.splitlines()[0]
    assert "context -8/+8" in txt


def test_model_loader_errors(lib, tiny_model, tmp_path):
    with pytest.raises(lib.RsError, match="cannot open"):
        lib.model_check(str(tmp_path / "missing.mdl"), tiny_model.online_conf)
    bad = tmp_path / "online.conf"
    bad.write_text("--feature-type=plp\n")
    with pytest.raises(lib.RsError, match="feature-type"):
        lib.model_check(tiny_model.final_mdl, str(bad))
    junk = tmp_path / "HCLG.fst"
    junk.write_bytes(b"not an fst at all, really")
    with pytest.raises(lib.RsError, match="magic"):
        lib.graph_check(str(junk), None)


def test_graph_loader(lib, tiny_model, synth, tmp_path):
    c = lib.graph_check(tiny_model.hclg, tiny_model.words_txt)
    assert c["states"] == tiny_model.num_states
    assert c["emitting_arcs"] + c["epsilon_arcs"] == tiny_model.num_arcs
    assert c["epsilon_arcs"] > 0 and c["final_states"] > 0 and c["start"] == 0
    assert c["words"] == len(tiny_model.words) + 2      # <eps> ... #0
    # the aligned ConstFst variant (version 1) reads identically
    p2 = synth.write_model(str(tmp_path), synth.TINY, aligned_fst=True)
    assert lib.graph_check(p2.hclg, p2.words_txt) == c


def test_python_surface_mirrors_reference():
    """Same constructor / method signatures as reference transcribe_wav.py:16-42, transcribe_stream.py:19-45."""
    import rhasspy_speech_b200 as pkg
    for cls in (pkg.KaldiNnet3WavTranscriber, pkg.KaldiNnet3StreamTranscriber):
        params = list(inspect.signature(cls.__init__).parameters)
        assert params[:8] == ["self", "model_dir", "graph_dir", "tools", "max_active", "lattice_beam", "acoustic_scale", "beam"]
        d = inspect.signature(cls.__init__).parameters
        assert (d["max_active"].default, d["lattice_beam"].default, d["acoustic_scale"].default, d["beam"].default) == (7000, 8.0, 1.0, 24.0)
        tp = list(inspect.signature(cls.async_transcribe).parameters)
        assert tp[2:] == ["lang_dir", "nbest", "max_fuzzy_cost", "require_fuzzy"]
        assert inspect.iscoroutinefunction(cls.async_transcribe) and inspect.iscoroutinefunction(cls.async_transcribe_rescore)
    legacy = list(inspect.signature(pkg.KaldiTranscriber.__init__).parameters)
    assert legacy[:4] == ["self", "model_dir", "graph_dir", "kaldi_bin_dir"]
    assert hasattr(pkg.KaldiTranscriber, "transcribe_wav") and hasattr(pkg.KaldiTranscriber, "transcribe_stream")


def test_decode_meta_and_nbest_text():
    import base64
    import json
    from rhasspy_speech_b200 import transcribe as T
    enc = lambda s: base64.b32encode(s.encode()).decode()
    assert T.decode_meta("turn on the light") == "turn on the light"
    word = T.OUTPUT_PREFIX + enc(json.dumps({"text": "kitchen", "list": "area"}))
    sent = T.SENTENCE_OUTPUT + enc("lights on in {area}")
    assert T.decode_meta("turn on " + word) == "turn on kitchen"
    assert T.decode_meta("turn on " + word + " " + sent) == "lights on in kitchen"
    # the fast path (no "__" in the text) must not swallow look-alikes; a lone marker prefix without payload is text
    assert T.decode_meta("under_score and double__underscore stay") == "under_score and double__underscore stay"
    assert T.decode_meta("set " + word + " " + word) == "set kitchen kitchen"

    class H:
        words = [[12, 45, 7], None, [], [3, 4]]
        nbest = [[([12, 45, 7], 1.0, 2.0)], [], [([], 0.0, 0.0)], [([3, 4], 1.0, 2.0), ([3], 1.5, 2.5), ([], 9.0, 1.0)]]
    assert T.nbest_text(H, 0) == b"utt-1 12 45 7 \n"       # kaldi-holder-inl.h:244-251: each int is followed by a space
    assert T.nbest_text(H, 1) == b""
    assert T.nbest_text(H, 2) == b"utt-1 \n"
    assert T.nbest_text(H, 3) == b"utt-1 3 4 \nutt-2 3 \nutt-3 \n"   # lattice-to-nbest.cc:104-107 keys


def test_python_copy_of_a_result_block():
    """Hypotheses(rs_result): the one-best and the n-best views of a fabricated result block (plain ints and floats out,
    utterances without a hypothesis as None / [], n-best lists in the order the library wrote them)."""
    import ctypes as C
    from rhasspy_speech_b200 import _lib

    def ptr(a):
        return a.ctypes.data_as(C.POINTER(C.c_int32 if a.dtype == np.int32 else C.c_float))
    n_hyp = np.array([1, 0, 1], np.int32)
    off = np.array([0, 3, 3, 5], np.int32)
    ids = np.array([12, 45, 7, 3, 4], np.int32)
    gc, ac = np.array([1.5, 0.0, 2.25], np.float32), np.array([-10.0, 0.0, -20.5], np.float32)
    frames, status = np.array([100, 90, 80], np.int32), np.array([0, 4, 64], np.int32)
    r = _lib.Result(3, ptr(n_hyp), ptr(off), ptr(ids), ptr(gc), ptr(ac), ptr(frames), ptr(status))
    h = _lib.Hypotheses(r)
    assert h.words == [[12, 45, 7], None, [3, 4]] and all(type(x) is int for w in h.words if w for x in w)
    assert h.nbest == [[([12, 45, 7], 1.5, -10.0)], [], [([3, 4], 2.25, -20.5)]]
    assert list(h.status) == [0, 4, 64] and list(h.num_frames) == [100, 90, 80]
    # n-best block: utterance 0 has three hypotheses, utterance 2 two; the first of each is the one-best
    n_hyp = np.array([3, 0, 2], np.int32)
    ho = np.array([0, 3, 3, 5], np.int32)
    wo = np.array([0, 3, 5, 5, 7, 8], np.int32)
    wid = np.array([12, 45, 7, 12, 45, 3, 4, 3], np.int32)
    hg, ha = np.arange(5, dtype=np.float32), -np.arange(5, dtype=np.float32)
    r = _lib.Result(3, ptr(n_hyp), ptr(off), ptr(ids), ptr(gc), ptr(ac), ptr(frames), ptr(status), ptr(ho), ptr(wo), ptr(wid), ptr(hg), ptr(ha))
    h = _lib.Hypotheses(r)
    assert h.nbest[0] == [([12, 45, 7], 0.0, -0.0), ([12, 45], 1.0, -1.0), ([], 2.0, -2.0)]
    assert h.nbest[1] == [] and h.nbest[2] == [([3, 4], 3.0, -3.0), ([3], 4.0, -4.0)]
    assert all(type(c) is float for u in h.nbest for _, g, a in u for c in (g, a))


def test_shard_utterances_balances_audio():
    from rhasspy_speech_b200.shard import shard_utterances
    rng = np.random.default_rng(3)
    dur = rng.uniform(0.8, 5.0, 2048).tolist()
    for world in (1, 2, 4, 8):
        parts = shard_utterances(dur, world)
        assert sorted(i for p in parts for i in p) == list(range(2048))
        loads = [sum(dur[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= 5.0
    assert shard_utterances([], 4) == [[], [], [], []]


def test_dynamic_batcher_groups_concurrent_requests():
    """transcribe._Batcher: requests that arrive while a batch is running share the next device batch, grouped
    by their n-best setting; a failing batch is retried per request so only the bad call raises."""
    import threading
    import time
    from types import SimpleNamespace
    from rhasspy_speech_b200 import transcribe as T

    class FakeDecoder:
        def __init__(self):
            self.calls = []
            self.nbest = None
            self.gate = threading.Event()

        def set_nbest(self, n, scale):
            self.nbest = (n, scale)

        def decode_wavs(self, paths):
            self.gate.wait(5)
            self.calls.append((self.nbest, list(paths)))
            if any("bad" in p for p in paths):
                raise T._lib.RsError("cannot open bad.wav")
            n = len(paths)
            return SimpleNamespace(words=[[len(p)] for p in paths], nbest=[[([len(p)], 0.0, 0.0)] for p in paths],
                                   status=[0] * n, n_hyp=[1] * n)

    dec = FakeDecoder()
    b = T._Batcher(dec, threading.Lock(), max_batch=4)
    first = b.submit("wav", "a.wav", 1, 1.0)            # starts running, blocks on the gate
    time.sleep(0.05)
    rest = [b.submit("wav", "u%d.wav" % i, 1, 1.0) for i in range(6)]
    other = b.submit("wav", "n5.wav", 5, 1.0)           # different n-best: its own batch
    bad = [b.submit("wav", name, 1, 0.5) for name in ("ok1.wav", "bad.wav", "ok22.wav")]
    dec.gate.set()
    assert first.result(5).words == [[5]]
    assert [f.result(5).words[0] for f in rest] == [[6]] * 6
    assert other.result(5).nbest == [[([6], 0.0, 0.0)]]
    assert bad[0].result(5).words == [[7]] and bad[2].result(5).words == [[8]]
    with pytest.raises(T._lib.RsError):
        bad[1].result(5)
    sizes = [len(p) for _, p in dec.calls]
    assert sizes[:3] == [1, 4, 2]                       # lone request, then the burst in max_batch pieces
    assert ((5, 1.0), ["n5.wav"]) in dec.calls
    assert ((1, 0.5), ["ok1.wav", "bad.wav", "ok22.wav"]) in dec.calls     # the batch that failed ...
    assert ((1, 0.5), ["bad.wav"]) in dec.calls                            # ... and its per-request retry
    assert b.batches[:3] == [1, 4, 2]


def test_transcriber_flow_with_a_fake_engine(tmp_path, monkeypatch):
    """The Python mirror end to end on CPU with the C library replaced by a fake decoder: async_transcribe (through
    the dynamic batcher), async_transcribe_many, the legacy class, n-best text and word mapping."""
    import asyncio
    import threading
    from types import SimpleNamespace
    from rhasspy_speech_b200 import transcribe as T

    words = {1: "turn", 2: "on", 3: "off", 4: "light"}

    class FakeGraph:
        def word(self, i):
            return words.get(i)

    class FakeDecoder:
        graph = FakeGraph()

        def __init__(self):
            self.nbest = 1

        def set_nbest(self, n, scale=1.0):
            self.nbest = n

        def decode_wavs(self, paths):
            n = len(paths)
            lists = [[([1, 2, 4], 1.0, 2.0), ([1, 3, 4], 1.5, 2.5)][:self.nbest] for _ in paths]
            return SimpleNamespace(n_utts=n, words=[l[0][0] for l in lists], nbest=lists, status=[0] * n, n_hyp=[len(l) for l in lists])

    dec = FakeDecoder()
    lock = threading.Lock()
    eng = SimpleNamespace(decoder=dec, lock=lock, graph=dec.graph, batcher=T._Batcher(dec, lock),
                          words=lambda ids, graph=None: " ".join((graph or dec.graph).word(i) for i in ids))
    monkeypatch.setattr(T._Base, "_get_engines", lambda self: [eng])
    monkeypatch.setattr(T.KaldiTranscriber, "_get_engine", lambda self: eng)
    tr = T.KaldiNnet3WavTranscriber(tmp_path, tmp_path, None)
    assert asyncio.run(tr.async_transcribe("a.wav", tmp_path)) == ["turn on light"]
    assert asyncio.run(tr.async_transcribe("a.wav", tmp_path, nbest=2)) == ["turn on light", "turn off light"]
    many = asyncio.run(tr.async_transcribe_many(["a.wav", "b.wav"], tmp_path, nbest=2))
    assert many == [["turn on light", "turn off light"]] * 2
    assert T.KaldiTranscriber(tmp_path, tmp_path).transcribe_wav("a.wav") == "turn on light"
    with pytest.raises(RuntimeError):
        asyncio.run(tr.async_transcribe("a.wav", tmp_path, nbest=0))
    with pytest.raises(NotImplementedError):
        asyncio.run(tr.async_transcribe_rescore("a.wav", tmp_path, tmp_path))


def test_cancelled_requests_do_not_kill_the_batcher(tmp_path):
    """A caller that gives up (asyncio.wait_for timeout, client disconnect) cancels its future while the request is queued
    or running.  The worker thread must survive, the co-batched callers must still get their results, and a stream handle
    is closed by the batcher -- after the decode, never under it."""
    import asyncio
    import threading
    import time
    from types import SimpleNamespace
    from rhasspy_speech_b200 import transcribe as T

    class FakeStream:
        def __init__(self, name):
            self.name, self.closed_at = name, None

        def accept(self, chunk):
            pass

        def close(self):
            self.closed_at = time.monotonic()

    class FakeDecoder:
        def __init__(self):
            self.gate = threading.Event()
            self.calls = []
            self.done_at = None

        def set_nbest(self, n, scale):
            pass

        def finish_streams(self, streams):
            self.gate.wait(5)
            assert all(s.closed_at is None for s in streams)      # nobody freed a stream that is being decoded
            self.calls.append([s.name for s in streams])
            self.done_at = time.monotonic()
            n = len(streams)
            return SimpleNamespace(words=[[len(s.name)] for s in streams], nbest=[[([len(s.name)], 0.0, 0.0)] for s in streams],
                                   status=[0] * n, n_hyp=[1] * n)

    dec = FakeDecoder()
    b = T._Batcher(dec, threading.Lock())
    s0, s1, s2, s3 = (FakeStream(n) for n in ("a", "bb", "ccc", "dddd"))
    running = b.submit("stream", s0, 1, 1.0)          # picked up at once, blocks on the gate
    time.sleep(0.05)
    queued = [b.submit("stream", s, 1, 1.0) for s in (s1, s2, s3)]
    assert queued[1].cancel()                          # cancelled while queued
    assert not running.cancel()                        # already running: cannot be cancelled under the decode
    dec.gate.set()
    assert running.result(5).words == [[1]]
    assert queued[0].result(5).words == [[2]] and queued[2].result(5).words == [[4]]
    assert dec.calls == [["a"], ["bb", "dddd"]]        # the cancelled request never reached the device batch
    assert b.thread is not None and b.thread.is_alive()
    for s in (s0, s1, s2, s3):
        assert s.closed_at is not None                  # every handle released by the batcher ...
    assert s0.closed_at >= dec.done_at - 1.0
    # ... and the worker still serves new requests
    again = b.submit("stream", FakeStream("ee"), 1, 1.0)
    assert again.result(5).words == [[2]]

    # the coroutine-level view: wait_for times out on one of two concurrent calls, the other completes
    dec2 = FakeDecoder()
    lock = threading.Lock()
    eng = SimpleNamespace(decoder=SimpleNamespace(open_stream=lambda: FakeStream("s"), graph=None), lock=lock, graph=None,
                          batcher=T._Batcher(dec2, lock), words=lambda ids, graph=None: "w%d" % ids[0])
    tr = T.KaldiNnet3StreamTranscriber(tmp_path, tmp_path, None)
    tr._get_engine = lambda: eng

    async def chunks():
        yield b"\\0\\0"

    async def main():
        slow = asyncio.ensure_future(tr.async_transcribe(chunks(), tmp_path))
        fast = asyncio.ensure_future(asyncio.wait_for(tr.async_transcribe(chunks(), tmp_path), timeout=0.2))
        await asyncio.sleep(0.4)
        dec2.gate.set()
        out = await slow
        with pytest.raises(asyncio.TimeoutError):
            await fast
        return out
    assert asyncio.run(main()) == ["w1"]


def test_capacity_status_bits_raise_like_a_failing_kaldi_binary():
    from rhasspy_speech_b200 import transcribe as T
    for bit in (1, 2, 8):
        with pytest.raises(RuntimeError, match="Unexpected error running command"):
            T.check_status(bit, "online2-wav-nnet3-latgen-faster")
    for ok in (0, 4, 16, 16 | 64, 32, 16 | 256 | 1024):
        T.check_status(ok, "online2-wav-nnet3-latgen-faster")


def test_compressed_matrices_expand_as_the_reference_does(tmp_path):
    """CM / CM2 / CM3 files written by the reference's copy-feats (tests/golden/cm_golden.npz, make_cm_golden.py) through
    the library's reader: the values the reference itself prints for them (6 significant digits in text mode)."""
    import __graft_entry__ as g
    g.build()
    from rhasspy_speech_b200 import _lib
    from conftest import golden_dir
    gold = np.load(os.path.join(golden_dir(), "cm_golden.npz"))
    for tok in ("CM", "CM2", "CM3"):
        f = tmp_path / (tok + ".mat")
        f.write_bytes(gold[tok + "_file"].tobytes())
        got = _lib.read_matrix(str(f))
        want = gold[tok + "_expanded"]
        assert got.shape == want.shape
        assert np.allclose(got, want, rtol=2e-6, atol=1e-6), (tok, np.abs(got - want).max())
        # and the compression did lose precision against the source, i.e. the expansion was really exercised
        assert np.abs(got - gold["source"]).max() > 1e-4


def test_multi_device_pool_deals_requests_over_engines(tmp_path, monkeypatch):
    """SURVEY 8e on the product path: a transcriber given several devices keeps one engine per device; a request list is
    dealt longest-first (shard.shard_utterances) and the shares run concurrently; single requests and streams go to the
    least-loaded engine.  Fake engines on CPU; tests/test_gpu_surface.py runs the same path on real devices."""
    import asyncio
    import threading
    import time
    from types import SimpleNamespace
    from rhasspy_speech_b200 import transcribe as T

    class FakeGraph:
        def word(self, i):
            return "w%d" % i

    class FakeDecoder:
        def __init__(self, name):
            self.name, self.graph, self.calls, self.busy = name, FakeGraph(), [], threading.Event()

        def set_nbest(self, n, scale=1.0):
            pass

        def decode_wavs(self, paths):
            self.calls.append(list(paths))
            time.sleep(0.05)
            n = len(paths)
            ids = [[int(os.path.basename(p)[1:4])] for p in paths]
            return SimpleNamespace(n_utts=n, words=ids, nbest=[[(w, 0.0, 0.0)] for w in ids], status=[0] * n, n_hyp=[1] * n)

    engines = []
    for k in range(3):
        dec = FakeDecoder("gpu%d" % k)
        lock = threading.Lock()
        engines.append(SimpleNamespace(decoder=dec, lock=lock, graph=dec.graph, batcher=T._Batcher(dec, lock), open_streams=0, device=k,
                                       words=lambda ids, graph=None: " ".join("w%d" % i for i in ids)))
    monkeypatch.setattr(T._Base, "_get_engines", lambda self: engines)
    monkeypatch.setattr(T, "_SHARE_TARGET_BYTES", 8000.0)               # ~ a third of the list below: it fills three devices
    tr = T.KaldiNnet3WavTranscriber(tmp_path, tmp_path, None, device=[0, 1, 2])
    paths = []
    for i in range(20):
        p = tmp_path / ("u%03d.wav" % i)
        p.write_bytes(b"x" * (1000 + 977 * ((i * 7) % 11)))
        paths.append(str(p))
    t0 = time.monotonic()
    out = asyncio.run(tr.async_transcribe_many(paths, tmp_path))
    assert out == [["w%d" % i] for i in range(20)]                       # original order restored
    shares = [e.decoder.calls[0] for e in engines]
    assert sorted(p for s in shares for p in s) == sorted(paths) and all(len(s) >= 5 for s in shares)
    loads = [sum(os.path.getsize(p) for p in s) for s in shares]
    assert max(loads) - min(loads) <= 11 * 977                           # equal audio per device (longest-first deal)
    assert time.monotonic() - t0 < 0.14                                  # the three shares ran side by side, not 3 x 50 ms
    assert not any(T._DEVICE_BYTES.get(k, 0.0) for k in range(3))        # the booking is released
    # a list too small to fill more than one device stays whole and goes to the least-loaded device: with several
    # (model, graph) pairs active, every device then runs full batches instead of every pair splitting its list
    monkeypatch.setattr(T, "_SHARE_TARGET_BYTES", 1.0e9)
    for e in engines:
        e.decoder.calls.clear()
    async def two_lists():
        return await asyncio.gather(tr.async_transcribe_many(paths[:10], tmp_path), tr.async_transcribe_many(paths[10:], tmp_path))
    a, b = asyncio.run(two_lists())
    assert a == [["w%d" % i] for i in range(10)] and b == [["w%d" % i] for i in range(10, 20)]
    used = [e.decoder.calls for e in engines if e.decoder.calls]
    assert len(used) == 2 and sorted(len(c[0]) for c in used) == [10, 10]   # two whole lists on two different devices
    plan = T.plan_shares([5.0, 1.0, 4.0, 2.0], [7, 8, 9], {7: 100.0, 8: 0.0, 9: 3.0}, 6.0)
    assert plan == [(1, [0, 1]), (2, [2, 3])]                            # two shares of 6 on the two least-loaded devices
    # single requests spread over the pool
    async def burst():
        return await asyncio.gather(*[tr.async_transcribe(paths[i], tmp_path) for i in range(9)])
    for e in engines:
        e.decoder.calls.clear()
    assert asyncio.run(burst()) == [["w%d" % i] for i in range(9)]
    assert all(e.decoder.calls for e in engines)
    assert T.resolve_devices(3) == [3] and T.resolve_devices([1, 1, 0]) == [1, 1, 0]
    with pytest.raises(ValueError):
        T.resolve_devices("some")
