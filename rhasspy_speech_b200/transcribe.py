"""Python surface of the decoder: drop-in for rhasspy-speech's transcribers on the hot path.

Mirrors, with the same names, argument meaning and error behaviour:

* ``rhasspy_speech.transcribe_wav.KaldiNnet3WavTranscriber``      (reference transcribe_wav.py:15-105)
* ``rhasspy_speech.transcribe_stream.KaldiNnet3StreamTranscriber``  (reference transcribe_stream.py:18-129)
* the legacy ``rhasspy_speech.KaldiTranscriber`` spelling used by the reference's own tests
  (tests/test_en_US-zamia.py:7,46-57).

Where the reference spawns ``online2-wav-nnet3-latgen-faster | lattice-to-nbest | nbest-to-linear``
(or ``online2-cli-nnet3-decode-faster``) per call, these classes call the C ABI of librs_b200.so once;
models and graphs are loaded once per (model_dir, graph_dir) and stay resident on the GPU.  The
intermediate the reference passes around -- ``nbest_stdout``, a Kaldi text Int32Vector archive with
one ``utt-<k> id id ... \\n`` line per hypothesis (kaldi/src/util/kaldi-holder-inl.h:244-251,
latbin/lattice-to-nbest.cc:104-107) -- is reproduced byte for byte so the unchanged tail of the
reference (int2sym, ``get_fuzzy_text``, ``decode_meta``) can consume it.

There is no CPU fallback: without the CUDA library or a GPU these classes raise.
"""
from __future__ import annotations

import asyncio
import base64
import collections
import concurrent.futures
import json
import logging
import os
import re
import threading
from pathlib import Path
from typing import AsyncIterable, Dict, Iterable, List, Optional, Sequence, Tuple, Union

from . import _lib

OUTPUT_PREFIX = "__output:"             # reference hassil_fst.py:32-33
SENTENCE_OUTPUT = "__sentence_output:"

_ENGINES: Dict[Tuple, "_Engine"] = {}
_ENGINES_LOCK = threading.Lock()


def decode_meta_single(text: str) -> str:
    return base64.b32decode(text.encode("utf-8")).strip().decode("utf-8")


_OUTPUT_RE = re.compile(re.escape(OUTPUT_PREFIX) + "([0-9A-Z=]+)")
_SENTENCE_RE = re.compile(re.escape(SENTENCE_OUTPUT) + "([0-9A-Z=]+)")


def decode_meta(text: str) -> str:
    """Restatement of reference hassil_fst.py:849-868 (output words carry base32 JSON metadata)."""
    if "__" not in text:        # neither marker can match (both start with two underscores): the common transcript
        return text
    slots: Dict[str, str] = {}

    def handle_match(m: "re.Match") -> str:
        data = json.loads(decode_meta_single(m.group(1)))
        slot_name = data.get("list")
        slot_value = data["text"]
        if slot_name:
            slots[slot_name] = slot_value
        return slot_value

    text = _OUTPUT_RE.sub(handle_match, text)
    match = _SENTENCE_RE.search(text)
    if match is None:
        return text
    return decode_meta_single(match.group(1)).format(**slots)


STATUS_ERRORS = ((1, "more tokens on one frame than max_tokens_per_frame"),
                 (2, "more tokens in the utterance than the traceback arena holds (max_tokens_per_utt)"),
                 (8, "more words in the transcript than max_words"))
_log = logging.getLogger(__name__)


def check_status(status: int, command: str):
    """rs_result.status of one utterance: capacity errors (bits 0, 1, 3) would otherwise show up as an empty or truncated
    transcript where the reference produces one; they are raised like a failing Kaldi binary (tools.py:81-88).  Bit 2 (no
    surviving tokens) is the reference's empty lattice: no hypothesis, not an error.  Bits 4-6 are information."""
    for bit, what in STATUS_ERRORS:
        if status & bit:
            raise RuntimeError("Unexpected error running command %s: decoder capacity exceeded: %s" % (command, what))
    if status & 32:
        _log.warning("n-best requested but the lattice did not fit its device buffers (RS_B200_LATTICE_MAX_MB): best path only")
    if status & 64:
        _log.debug("utterance was order-sensitive on the device and was decoded by the strict-order host decoder")


class _OneHyp:
    """The slice of a batch result that belongs to one request: what nbest_text() reads, for utterance 0."""

    def __init__(self, hyp, u: int, graph=None):
        self.graph = graph              # the graph this request was decoded on (its words.txt names the ids)
        self.words = [hyp.words[u]]
        self.nbest = [hyp.nbest[u]]
        self.status = [int(hyp.status[u])]
        self.n_hyp = [int(hyp.n_hyp[u])]


class _Batcher:
    """Host dynamic batcher (threading): concurrent requests on one engine become ONE device batch.

    The reference runs one OS process per call, so 64 concurrent streams are 64 processes.  Here a request only
    enqueues (kind, payload, nbest, ranking scale) and gets a future; one worker thread per engine drains the
    queue: every request that arrived while the previous batch was on the GPU, and shares its n-best setting, goes
    into the next rs_decode_wavs / rs_streams_finish call.  No linger by default (RS_B200_BATCH_LINGER_MS): a lone
    request is decoded at once, a burst batches itself behind the request that is running.  A failing batch is
    retried request by request so that only the offending call raises, as with the reference's independent
    processes."""

    def __init__(self, decoder, lock: threading.Lock, max_batch: int = 256):
        self.decoder, self.lock, self.max_batch = decoder, lock, max_batch
        self.linger = float(os.environ.get("RS_B200_BATCH_LINGER_MS", "0")) / 1e3
        self.cv = threading.Condition()
        self.pending: "collections.deque" = collections.deque()
        self.batches: List[int] = []          # size of every device batch so far (instrumentation)
        self.running = 0                      # requests of the batch that is on the device right now
        self.thread: Optional[threading.Thread] = None

    def submit(self, kind: str, payload, nbest: int, scale: float) -> "concurrent.futures.Future":
        fut: "concurrent.futures.Future" = concurrent.futures.Future()
        with self.cv:
            self.pending.append((kind, payload, int(nbest), float(scale), fut))
            if self.thread is None or not self.thread.is_alive():
                self.thread = threading.Thread(target=self._run, name="rs-b200-batcher", daemon=True)
                self.thread.start()
            self.cv.notify()
        return fut

    def _take(self):
        """Next batch: the oldest request and every queued request with the same (kind, nbest, scale)."""
        with self.cv:
            while not self.pending:
                if not self.cv.wait(timeout=1.0) and not self.pending:
                    self.thread = None          # idle: let the thread end; submit() starts a new one
                    return None
            if self.linger > 0:
                self.cv.wait(timeout=self.linger)
            key = self.pending[0][:1] + self.pending[0][2:4]
            batch, rest = [], collections.deque()
            while self.pending:
                item = self.pending.popleft()
                if len(batch) < self.max_batch and item[:1] + item[2:4] == key:
                    batch.append(item)
                else:
                    rest.append(item)
            self.pending = rest
            return batch

    def _decode(self, kind: str, payloads, nbest: int, scale: float):
        with self.lock:
            self.decoder.set_nbest(nbest, scale)
            self.batches.append(len(payloads))
            if kind == "wav":
                hyp = self.decoder.decode_wavs([str(p) for p in payloads])
            elif kind == "stream":
                hyp = self.decoder.finish_streams(payloads)
            else:
                hyp = self.decoder.decode_pcm(payloads)
            hyp.graph = getattr(self.decoder, "graph", None)
            return hyp

    @staticmethod
    def _deliver(fut: "concurrent.futures.Future", result=None, error: Optional[BaseException] = None):
        """Resolve one request's future; a future that was cancelled meanwhile (asyncio.wait_for timeout, client gone)
        must not take the worker thread -- and with it every other caller of the engine -- down."""
        try:
            if error is not None:
                fut.set_exception(error)
            else:
                fut.set_result(result)
        except concurrent.futures.InvalidStateError:
            pass

    def _release(self, kind: str, payload):
        """The batcher owns a stream handle from submit() on: it is closed here, after the decode (or when the request was
        cancelled before it ran), never by the awaiting coroutine -- so a cancelled caller cannot free audio in flight."""
        if kind == "stream":
            try:
                payload.close()
            except Exception:  # noqa: BLE001
                pass

    def _run(self):
        while True:
            batch = self._take()
            if batch is None:
                return
            # a request whose caller was cancelled while it was queued is dropped here (set_running_or_notify_cancel
            # returns False), the others are marked running and can no longer be cancelled under the decode
            live = []
            for b in batch:
                if b[4].set_running_or_notify_cancel():
                    live.append(b)
                else:
                    self._release(b[0], b[1])
            batch = live
            if not batch:
                continue
            kind, _, nbest, scale, _ = batch[0]
            self.running = len(batch)
            try:
                try:
                    hyp = self._decode(kind, [b[1] for b in batch], nbest, scale)
                    for u, b in enumerate(batch):
                        self._deliver(b[4], _OneHyp(hyp, u, hyp.graph))
                except Exception as first:  # noqa: BLE001 -- delivered to the caller(s) below
                    if len(batch) == 1:
                        self._deliver(batch[0][4], error=first)
                        continue
                    for b in batch:             # isolate the offending request
                        try:
                            one = self._decode(kind, [b[1]], nbest, scale)
                            self._deliver(b[4], _OneHyp(one, 0, one.graph))
                        except Exception as e:  # noqa: BLE001
                            self._deliver(b[4], error=e)
            finally:
                self.running = 0
                for b in batch:
                    self._release(b[0], b[1])


class _Engine:
    """One resident (model, graph, decoder) triple; serialises calls (an rs_decoder is single-threaded)."""

    def __init__(self, final_mdl: Path, online_conf: Path, hclg: Path, words_txt: Path, device: int, **opts):
        for f in (final_mdl, online_conf, hclg, words_txt):
            if not Path(f).is_file():
                # the reference fails with the Kaldi binary's stderr; keep its exception type and prefix
                raise RuntimeError("Unexpected error running command online2-wav-nnet3-latgen-faster: cannot open %s" % f)
        self.device = device
        self.graph_files = (Path(hclg), Path(words_txt))
        self.model_sig = _file_sig((final_mdl, online_conf))
        self.graph_sig = _file_sig(self.graph_files)
        self.model = _lib.Model(str(final_mdl), str(online_conf), device)
        self.graph = _lib.Graph(str(hclg), str(words_txt), device)
        self.decoder = _lib.Decoder(self.model, self.graph, **opts)
        self.lock = threading.Lock()
        self.batcher = _Batcher(self.decoder, self.lock)
        self.open_streams = 0           # streams bound to this engine (sticky stream -> GPU assignment)

    def refresh_graph(self):
        """Hot graph swap.  The reference reads HCLG.fst and words.txt in every call, so the graph
        KaldiTrainer writes (kaldi.py:409-425) is used by the next transcription; the resident engine gets the same
        behaviour by re-binding its decoder when the files changed (rs_graph_load + rs_decoder_set_graph: the model
        and the decoder's device workspace stay)."""
        sig = _file_sig(self.graph_files)
        if sig == self.graph_sig:
            return
        with self.lock:
            if sig == self.graph_sig:
                return
            try:
                new = _lib.Graph(str(self.graph_files[0]), str(self.graph_files[1]), self.device)
                self.decoder.set_graph(new)
            except _lib.RsError as e:
                raise RuntimeError("Unexpected error running command online2-wav-nnet3-latgen-faster: %s" % e) from e
            self.graph = new            # the old graph is freed when the last result decoded on it is gone
            self.graph_sig = sig

    def words(self, ids: Sequence[int], graph=None) -> str:
        out = []
        for i in ids:
            w = (graph or self.graph).word(i)
            if w is None:
                raise RuntimeError("Unexpected error running command int2sym.pl: undefined symbol %d" % i)
            out.append(w)
        return " ".join(out)


# Audio bytes outstanding per DEVICE, over all transcribers of the process: several (model, graph) pairs share the
# devices, and a request list is only split over as many devices as it can fill.
_DEVICE_BYTES: Dict[int, float] = {}
_DEVICE_BYTES_LOCK = threading.Lock()
# a share smaller than this no longer fills a GPU (8 MB of 16-bit 16 kHz PCM ~ 64 utterances of 4 s): shorter lists go
# to fewer devices.  (Measured on 2 GPUs, 8 pairs x 256 short utterances = 16 MB each: split over both devices 37.9 ms,
# whole lists on the less loaded device 45.5 ms per 2048 utterances -- run-to-run spread of the concurrent calls is
# larger than the difference, scripts/debug_pool.py.)
_SHARE_TARGET_BYTES = float(os.environ.get("RS_B200_SHARE_TARGET_MB", "8")) * (1 << 20)


def plan_shares(sizes: Sequence[float], devices: Sequence[int], outstanding: Dict[int, float],
                target: float) -> List[Tuple[int, List[int]]]:
    """(position in `devices`, utterance indices) per share: the list goes to the k least-loaded devices, k = as many
    as it can fill with `target` bytes each, dealt longest-first (shard.shard_utterances)."""
    from .shard import shard_utterances
    total = float(sum(sizes))
    k = int(max(1, min(len(devices), round(total / target) if target > 0 else len(devices))))
    order = sorted(range(len(devices)), key=lambda i: (outstanding.get(devices[i], 0.0), i))[:k]
    return [(order[j], idx) for j, idx in enumerate(shard_utterances(sizes, k))]


def _load(eng) -> int:
    b = getattr(eng, "batcher", None)
    return (len(b.pending) + getattr(b, "running", 0) if b is not None else 0) + getattr(eng, "open_streams", 0)


def _file_sig(paths) -> Tuple:
    """(mtime, size) of each file; missing files are left to the loader's error message."""
    out = []
    for p in paths:
        try:
            st = os.stat(p)
            out.append((st.st_mtime_ns, st.st_size))
        except OSError:
            out.append(None)
    return tuple(out)


_KEY_LOCKS: Dict[Tuple, threading.Lock] = {}


def _engine(final_mdl: Path, online_conf: Path, graph_dir: Path, device: int, max_active: int, beam: float,
            lattice_beam: float, replica: int = 0) -> _Engine:
    key = (str(final_mdl), str(online_conf), str(graph_dir), device, max_active, float(beam), float(lattice_beam), replica)
    with _ENGINES_LOCK:                 # held only for the dictionary lookups: loading one engine must not block the others
        key_lock = _KEY_LOCKS.setdefault(key, threading.Lock())
    with key_lock:
        with _ENGINES_LOCK:
            eng = _ENGINES.get(key)
        if eng is not None and eng.model_sig != _file_sig((final_mdl, online_conf)):
            eng = None                  # a new acoustic model: rebuild the engine (the old one is garbage-collected)
        if eng is not None:
            eng.refresh_graph()
        if eng is None:
            try:
                eng = _Engine(final_mdl, online_conf, graph_dir / "HCLG.fst", graph_dir / "words.txt", device,
                              max_active=max_active, beam=beam, lattice_beam=lattice_beam)
            except _lib.RsError as e:
                raise RuntimeError("Unexpected error running command online2-wav-nnet3-latgen-faster: %s" % e) from e
            with _ENGINES_LOCK:
                _ENGINES[key] = eng
        return eng


def nbest_text(hyp: "_lib.Hypotheses", utt: int) -> bytes:
    """The ``nbest-to-linear ... ark,t:-`` bytes for one utterance: ``utt-1 12 45 7 \\n`` (each id is
    followed by a space); empty when nothing was decoded, as when the reference's lattice is empty."""
    if hyp.words[utt] is None:
        return b""
    # utt-<k>: lattice-to-nbest appends "-<k>" to the utterance key (latbin/lattice-to-nbest.cc:104-107)
    return "".join("utt-%d %s\n" % (k + 1, "".join("%d " % w for w in words))
                   for k, (words, _, _) in enumerate(hyp.nbest[utt])).encode()


_FUZZY: Dict[str, Tuple[Tuple, "_lib.Fuzzy"]] = {}
_FUZZY_LOCK = threading.Lock()


def _fuzzy_matcher(lang_dir: Path) -> "_lib.Fuzzy":
    """lang_dir/G.fuzzy.fst + words.txt, loaded once and reloaded when training rewrote them."""
    files = (lang_dir / "G.fuzzy.fst", lang_dir / "words.txt")
    sig = _file_sig(files)
    with _FUZZY_LOCK:
        hit = _FUZZY.get(str(lang_dir))
        if hit is None or hit[0] != sig:
            try:
                hit = (sig, _lib.Fuzzy(str(files[0]), str(files[1])))
            except _lib.RsError as e:
                raise RuntimeError("Unexpected error running command fstcompose: %s" % e) from e
            _FUZZY[str(lang_dir)] = hit
        return hit[1]


async def _fuzzy(nbest_stdout: bytes, lang_dir: Path, tools) -> Optional[Tuple[str, float]]:
    """Out-of-vocabulary rejection (reference transcribe_util.py:11-88) in process: the hypotheses of
    ``nbest_stdout`` against lang_dir/G.fuzzy.fst (csrc/fuzzy.cc replaces fstcompile | fstcompose | fstshortestpath |
    fstrmepsilon | fsttopsort | fstproject | fstprint).  Without G.fuzzy.fst it is a no-op, as there.
    RS_B200_FUZZY=tools keeps the reference's own OpenFst pipeline (needs KaldiTools)."""
    if not (lang_dir / "G.fuzzy.fst").exists():
        return None
    if os.environ.get("RS_B200_FUZZY") == "tools":
        if tools is None or not hasattr(tools, "async_run_pipeline"):
            raise RuntimeError("Unexpected error running command fstcompile: G.fuzzy.fst is present but no KaldiTools were given")
        from rhasspy_speech.transcribe_util import get_fuzzy_text  # the unchanged reference tail
        return await get_fuzzy_text(nbest_stdout, lang_dir, tools)
    hyps = [[int(x) for x in line.split()[1:]] for line in nbest_stdout.decode("utf-8").splitlines() if line.strip()]
    if not hyps:
        return None
    fz = _fuzzy_matcher(lang_dir)
    try:
        hit = fz.match(hyps)
    except _lib.RsError as e:
        raise RuntimeError("Unexpected error running command fstcompose: %s" % e) from e
    if hit is None:
        return None
    words = [fz.word(i) or "" for i in hit[0]]
    words = [w for w in words if w and w != "<eps>"]
    return (" ".join(words), hit[1]) if words else None


def resolve_devices(device) -> List[int]:
    """`device` of the transcriber classes: one CUDA device index, a sequence of indices (a device may be listed more
    than once: that many engines on it), or "all"."""
    if isinstance(device, str):
        if device != "all":
            raise ValueError("device must be an index, a sequence of indices or 'all'")
        n = _lib.device_count()
        if n < 1:
            raise RuntimeError("Unexpected error running command online2-wav-nnet3-latgen-faster: no CUDA device available")
        return list(range(n))
    if isinstance(device, int):
        return [device]
    devs = [int(d) for d in device]
    if not devs:
        raise ValueError("empty device list")
    return devs


class _Base:
    def __init__(self, model_dir, graph_dir, tools=None, max_active: int = 7000, lattice_beam: float = 8.0,
                 acoustic_scale: float = 1.0, beam: float = 24.0, device: Union[int, Sequence[int], str] = 0):
        self.model_dir = Path(model_dir)
        self.graph_dir = Path(graph_dir)
        self.tools = tools
        self.max_active = max_active
        self.lattice_beam = lattice_beam
        # as in the reference, this scale is only applied when ranking n-best lists (lattice-to-nbest
        # --acoustic-scale, transcribe_wav.py:65); the search always runs at 1.0 (:54)
        self.acoustic_scale = acoustic_scale
        self.beam = beam
        # Multi-GPU: utterances are independent, so a transcriber given several devices keeps one engine
        # (model + graph replica, decoder, dynamic batcher) per device; a request list is dealt longest-first over them
        # (shard.shard_utterances), a single request or a stream goes to the least-loaded engine and stays there.
        self.device = device
        self._rr = 0

    def _paths(self) -> Tuple[Path, Path]:
        return (self.model_dir / "model" / "model" / "final.mdl", self.model_dir / "model" / "online" / "conf" / "online.conf")

    def _get_engines(self) -> List[_Engine]:
        final_mdl, online_conf = self._paths()
        devs = resolve_devices(self.device)
        seen: Dict[int, int] = {}
        jobs = []
        for d in devs:
            jobs.append((d, seen.get(d, 0)))
            seen[d] = seen.get(d, 0) + 1

        def make(job):
            return _engine(final_mdl, online_conf, self.graph_dir, job[0], self.max_active, self.beam, self.lattice_beam, job[1])
        if len(jobs) == 1:
            return [make(jobs[0])]
        with concurrent.futures.ThreadPoolExecutor(max_workers=len(jobs)) as ex:     # replicas load side by side
            return list(ex.map(make, jobs))

    def _get_engine(self) -> _Engine:
        """The engine for ONE request: the least-loaded of the pool (queued + running requests, then open streams)."""
        engines = self._get_engines()
        if len(engines) == 1:
            return engines[0]
        self._rr += 1
        return min(enumerate(engines), key=lambda ie: (_load(ie[1]), (ie[0] - self._rr) % len(engines)))[1]

    async def _get_engine_async(self) -> _Engine:
        """Engine lookup from a coroutine: the first call loads the model and the graph (disk reads, H2D upload), a later
        one may re-bind a retrained graph -- both in a worker thread, so the event loop keeps serving the other requests."""
        return await asyncio.get_running_loop().run_in_executor(None, self._get_engine)

    async def _submit(self, eng: _Engine, kind: str, payload, nbest: int, command: str):
        """One request through the engine's dynamic batcher; concurrent callers share a device batch."""
        if nbest < 1:
            raise RuntimeError("Unexpected error running command lattice-to-nbest: --n must be >= 1")
        try:
            hyp = await asyncio.wrap_future(eng.batcher.submit(kind, payload, nbest, self.acoustic_scale))
        except _lib.RsError as e:
            raise RuntimeError("Unexpected error running command %s: %s" % (command, e)) from e
        check_status(hyp.status[0], command)
        return hyp

    def _set_nbest(self, eng: _Engine, nbest: int):
        """`lattice-to-nbest --n=<nbest> --acoustic-scale=<acoustic_scale>` (transcribe_wav.py:62-67); call with
        eng.lock held.  n = 1 at scale 1.0 is the device back-trace, anything else goes through the lattice."""
        if nbest < 1:
            raise RuntimeError("Unexpected error running command lattice-to-nbest: --n must be >= 1")
        eng.decoder.set_nbest(nbest, self.acoustic_scale)

    async def _finish(self, eng: _Engine, nbest_stdout: bytes, lang_dir, max_fuzzy_cost, require_fuzzy,
                      graph=None) -> List[str]:
        lang_dir = Path(lang_dir)
        fuzzy_result = await _fuzzy(nbest_stdout, lang_dir, self.tools)
        if fuzzy_result is not None:
            text, cost = fuzzy_result
            if cost <= max_fuzzy_cost:  # TypeError when max_fuzzy_cost is None, exactly as the reference
                return [decode_meta(text)]
        if require_fuzzy:
            return []
        texts: List[str] = []
        for line in nbest_stdout.decode().splitlines():
            if line.startswith("utt-"):
                parts = line.strip().split()
                if len(parts) > 1:      # the reference drops hypotheses without words (transcribe_wav.py:99-103)
                    texts.append(decode_meta(eng.words([int(x) for x in parts[1:]], graph)))
        return texts


class KaldiNnet3WavTranscriber(_Base):
    async def async_transcribe(self, wav_path, lang_dir, nbest: int = 1, max_fuzzy_cost: Optional[float] = None,
                               require_fuzzy: bool = False) -> List[str]:
        eng = await self._get_engine_async()
        hyp = await self._submit(eng, "wav", wav_path, nbest, "online2-wav-nnet3-latgen-faster")
        return await self._finish(eng, nbest_text(hyp, 0), lang_dir, max_fuzzy_cost, require_fuzzy, hyp.graph)

    async def async_transcribe_many(self, wav_paths: Sequence, lang_dir, nbest: int = 1,
                                    max_fuzzy_cost: Optional[float] = None,
                                    require_fuzzy: bool = False) -> List[List[str]]:
        """Batched extension: element i equals async_transcribe(wav_paths[i]).  One device batch per engine: with several
        devices the list is dealt longest-first (file size = duration) so that every GPU gets the same audio seconds, and
        the shares run concurrently (no collective -- the results are word ids gathered by this thread)."""
        loop = asyncio.get_running_loop()
        all_engines = await loop.run_in_executor(None, self._get_engines)
        paths = [str(p) for p in wav_paths]
        if len(all_engines) > 1:
            sizes = []
            for p in paths:
                try:
                    sizes.append(float(os.path.getsize(p)))
                except OSError:
                    sizes.append(0.0)           # a missing file fails in its own share with the loader's message
            # as many devices as the list can fill, the least-loaded ones first (other transcribers of the process --
            # other (model, graph) pairs -- share the devices: _DEVICE_BYTES counts the audio outstanding on each)
            with _DEVICE_BYTES_LOCK:
                plan = plan_shares(sizes, [e.device for e in all_engines], _DEVICE_BYTES, _SHARE_TARGET_BYTES)
                booked = [(all_engines[pos].device, float(sum(sizes[i] for i in idx))) for pos, idx in plan]
                for dev, b in booked:
                    _DEVICE_BYTES[dev] = _DEVICE_BYTES.get(dev, 0.0) + b
            engines = [all_engines[pos] for pos, _ in plan]
            shares = [idx for _, idx in plan]
        else:
            engines, shares, booked = all_engines, [list(range(len(paths)))], []

        def run(eng, idx):
            if not idx:
                return None, None
            with eng.lock:
                try:
                    self._set_nbest(eng, nbest)
                    return eng.decoder.decode_wavs([paths[i] for i in idx]), eng.decoder.graph
                except _lib.RsError as e:
                    raise RuntimeError("Unexpected error running command online2-wav-nnet3-latgen-faster: %s" % e) from e
        try:
            parts = await asyncio.gather(*[loop.run_in_executor(None, run, eng, idx) for eng, idx in zip(engines, shares)])
        finally:
            with _DEVICE_BYTES_LOCK:
                for dev, b in booked:
                    _DEVICE_BYTES[dev] = max(0.0, _DEVICE_BYTES.get(dev, 0.0) - b)
        out: List[Optional[List[str]]] = [None] * len(paths)
        fuzzy = (Path(lang_dir) / "G.fuzzy.fst").exists()       # one stat for the whole list
        for eng, idx, (hyp, graph) in zip(engines, shares, parts):
            for k, i in enumerate(idx):
                check_status(int(hyp.status[k]), "online2-wav-nnet3-latgen-faster")
                if fuzzy or require_fuzzy:
                    out[i] = await self._finish(eng, nbest_text(hyp, k), lang_dir, max_fuzzy_cost, require_fuzzy, graph)
                else:       # no fuzzy matcher in play: word ids -> text without the detour over the text archive
                    out[i] = [decode_meta(eng.words(words, graph)) for words, _, _ in (hyp.nbest[k] if hyp.words[k] is not None else [])
                              if words]
        return out  # type: ignore[return-value]

    async def async_transcribe_rescore(self, *args, **kwargs):
        raise NotImplementedError("lattice rescoring (reference transcribe_wav.py:107-232) is scope row f3")


class KaldiNnet3StreamTranscriber(_Base):
    async def async_transcribe(self, audio_stream: AsyncIterable[Optional[bytes]], lang_dir, nbest: int = 1,
                               max_fuzzy_cost: Optional[float] = None, require_fuzzy: bool = False) -> List[str]:
        """audio_stream yields raw 16 kHz mono s16le chunks of any size (reference transcribe_stream.py:38-82)."""
        eng = await self._get_engine_async()        # least-loaded engine; the stream stays on it (sticky)
        stream = eng.decoder.open_stream()
        eng.open_streams = getattr(eng, "open_streams", 0) + 1
        try:
            return await self._transcribe_on(eng, stream, audio_stream, lang_dir, nbest, max_fuzzy_cost, require_fuzzy)
        finally:
            eng.open_streams -= 1

    async def _transcribe_on(self, eng, stream, audio_stream, lang_dir, nbest, max_fuzzy_cost, require_fuzzy) -> List[str]:
        try:
            pending = b""
            async for chunk in audio_stream:
                if not chunk:
                    continue
                if not pending and not (len(chunk) & 1):
                    stream.accept(chunk)            # the common case: whole samples, no copy on this side
                    continue
                data = pending + chunk
                keep = len(data) & ~1
                pending = data[keep:]
                if keep:
                    stream.accept(data[:keep])
        except BaseException:
            stream.close()                          # not submitted yet: still ours
            raise
        # from here on the batcher owns the stream handle and closes it after the decode, also when this coroutine is
        # cancelled while the request is queued or in flight (a close here could free audio the GPU batch is reading)
        hyp = await self._submit(eng, "stream", stream, nbest, "online2-cli-nnet3-decode-faster")
        return await self._finish(eng, nbest_text(hyp, 0), lang_dir, max_fuzzy_cost, require_fuzzy, hyp.graph)

    async def async_transcribe_rescore(self, *args, **kwargs):
        raise NotImplementedError("lattice rescoring (reference transcribe_stream.py:131-243) is scope row f3")


class KaldiTranscriber:
    """Legacy synchronous spelling: ``KaldiTranscriber(model_dir, graph_dir, kaldi_bin_dir).transcribe_wav(path) -> str``
    (reference tests/test_en_US-zamia.py:46-57; there ``model_dir`` already points at ``<model>/model``)."""

    def __init__(self, model_dir, graph_dir, kaldi_bin_dir=None, max_active: int = 7000, lattice_beam: float = 8.0,
                 acoustic_scale: float = 1.0, beam: float = 24.0, device: int = 0):
        self.model_dir = Path(model_dir)
        self.graph_dir = Path(graph_dir)
        self.kaldi_bin_dir = kaldi_bin_dir
        self.max_active, self.lattice_beam, self.acoustic_scale, self.beam, self.device = max_active, lattice_beam, acoustic_scale, beam, device

    def _get_engine(self) -> _Engine:
        return _engine(self.model_dir / "model" / "final.mdl", self.model_dir / "online" / "conf" / "online.conf",
                       self.graph_dir, self.device, self.max_active, self.beam, self.lattice_beam)

    def _text(self, eng: _Engine, hyp, utt: int) -> str:
        return decode_meta(eng.words(hyp.words[utt])) if hyp.words[utt] else ""

    def transcribe_wav(self, wav_path) -> str:
        eng = self._get_engine()
        with eng.lock:
            try:
                eng.decoder.set_nbest(1, 1.0)
                hyp = eng.decoder.decode_wavs([str(wav_path)])
            except _lib.RsError as e:
                raise RuntimeError("Unexpected error running command online2-wav-nnet3-latgen-faster: %s" % e) from e
        check_status(int(hyp.status[0]), "online2-wav-nnet3-latgen-faster")
        return self._text(eng, hyp, 0)

    def transcribe_wavs(self, wav_paths: Sequence) -> List[str]:
        eng = self._get_engine()
        with eng.lock:
            eng.decoder.set_nbest(1, 1.0)
            hyp = eng.decoder.decode_wavs([str(p) for p in wav_paths])
        for u in range(len(wav_paths)):
            check_status(int(hyp.status[u]), "online2-wav-nnet3-latgen-faster")
        return [self._text(eng, hyp, u) for u in range(len(wav_paths))]

    def transcribe_stream(self, chunks: Iterable[bytes]) -> str:
        eng = self._get_engine()
        stream = eng.decoder.open_stream()
        try:
            pending = b""
            for chunk in chunks:
                if not chunk:
                    continue
                if not pending and not (len(chunk) & 1):
                    stream.accept(chunk)            # the common case: whole samples, no copy on this side
                    continue
                data = pending + chunk
                keep = len(data) & ~1
                pending = data[keep:]
                if keep:
                    stream.accept(data[:keep])
            with eng.lock:
                eng.decoder.set_nbest(1, 1.0)
                hyp = stream.finish()
        finally:
            stream.close()
        check_status(int(hyp.status[0]), "online2-cli-nnet3-decode-faster")
        return self._text(eng, hyp, 0)
