"""The Python mirror of the reference's transcribers under concurrency (BASELINE config 4, SURVEY 8b threading):
many coroutines on one engine share device batches through the host dynamic batcher, and every caller still gets
exactly the result of a lone call."""
import asyncio
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_concurrent_streams_and_wavs_share_device_batches(tiny_model, utterances, synth, tmp_path):
    import rhasspy_speech_b200 as pkg
    from rhasspy_speech_b200 import transcribe as T
    graph_dir = os.path.dirname(tiny_model.hclg)
    st = pkg.KaldiNnet3StreamTranscriber(tiny_model.model_dir, graph_dir, None)
    wt = pkg.KaldiNnet3WavTranscriber(tiny_model.model_dir, graph_dir, None)
    wavs = []
    for i, pcm in enumerate(utterances):
        w = os.path.join(str(tmp_path), "c%03d.wav" % i)
        synth.write_wav(w, pcm)
        wavs.append(w)

    async def chunks(pcm):
        raw = np.asarray(pcm, dtype="<i2").tobytes()
        for o in range(0, len(raw), 2560):          # 80 ms
            yield raw[o:o + 2560]
            await asyncio.sleep(0)                    # interleave the streams

    async def lone():
        out = []
        for pcm in utterances:
            out.append(await st.async_transcribe(chunks(pcm), tmp_path))
        return out, [await wt.async_transcribe(w, tmp_path, nbest=3) for w in wavs]

    async def burst(k):
        streams = [st.async_transcribe(chunks(utterances[i % len(utterances)]), tmp_path) for i in range(k)]
        files = [wt.async_transcribe(wavs[i % len(wavs)], tmp_path, nbest=3) for i in range(k)]
        return await asyncio.gather(*streams), await asyncio.gather(*files)

    want_s, want_w = asyncio.run(lone())
    assert all(len(x) == 1 for x in want_s) and any(len(x) > 1 for x in want_w)
    eng = st._get_engine()
    assert eng is wt._get_engine()                   # one resident (model, graph) pair for both surfaces
    before = len(eng.batcher.batches)
    k = 64
    got_s, got_w = asyncio.run(burst(k))
    assert got_s == [want_s[i % len(utterances)] for i in range(k)]
    assert got_w == [want_w[i % len(wavs)] for i in range(k)]
    sizes = eng.batcher.batches[before:]
    # (how the 128 requests fall into batches depends on the event loop's timing: the bounds leave room for a loaded box)
    assert sum(sizes) == 2 * k and len(sizes) <= 16 and max(sizes) >= k // 4, sizes
    # a request that cannot be decoded raises like the reference's failing process, and only for its caller
    async def mixed():
        return await asyncio.gather(wt.async_transcribe(wavs[0], tmp_path), wt.async_transcribe(str(tmp_path / "missing.wav"), tmp_path),
                                    return_exceptions=True)
    ok, err = asyncio.run(mixed())
    assert ok == want_w[0][:1] and isinstance(err, RuntimeError) and "online2-wav-nnet3-latgen-faster" in str(err)


def test_retrained_graph_is_picked_up_without_restart(ref, synth, utterances, tmp_path):
    """the reference reads HCLG.fst per call, so the graph KaldiTrainer rewrites is used by the next
    transcription.  The resident engine re-binds its decoder when the files change (rs_decoder_set_graph)."""
    import asyncio
    import dataclasses
    import shutil
    import rhasspy_speech_b200 as pkg
    a = synth.write_model(str(tmp_path / "a"), synth.TINY)
    b = synth.write_model(str(tmp_path / "b"), dataclasses.replace(synth.TINY, graph="arpa", vocab_size=120, bigrams_per_word=6, eps_hops=2))
    assert open(a.final_mdl, "rb").read() == open(b.final_mdl, "rb").read()       # same acoustic model, other graph
    wav = str(tmp_path / "x.wav")
    synth.write_wav(wav, utterances[2])

    def reference(p):
        want, _, _ = ref.transcribe_wavs(a.final_mdl, a.online_conf, p.hclg, p.words_txt, [wav])
        words = {int(l.split()[1]): l.split()[0] for l in open(p.words_txt)}
        return [" ".join(words[i] for i in want["utt00000-1"])] if want.get("utt00000-1") else []
    tr = pkg.KaldiNnet3WavTranscriber(a.model_dir, a.graph_dir, None)
    legacy = pkg.KaldiTranscriber(os.path.join(a.model_dir, "model"), a.graph_dir)
    first = asyncio.run(tr.async_transcribe(wav, tmp_path))
    assert first == reference(a) and first
    eng = tr._get_engine()
    n_states_a = eng.graph.num_states
    # "retrain": the graph files are replaced in place
    for f in ("HCLG.fst", "words.txt"):
        shutil.copyfile(os.path.join(b.graph_dir, f), os.path.join(a.graph_dir, f))
    second = asyncio.run(tr.async_transcribe(wav, tmp_path))
    assert second == reference(b) and second and second != first
    assert tr._get_engine() is eng and eng.graph.num_states != n_states_a          # same engine, new graph
    assert legacy.transcribe_wav(wav) == second[0]
    # n-best on the swapped graph
    many = asyncio.run(tr.async_transcribe(wav, tmp_path, nbest=3))
    assert many[:1] == second and len(many) >= 1


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_run
    if not ref_run.available():
        pytest.skip("oracle/_ref not built")
    return ref_run


def test_pinned_audio_block_gives_the_same_results(tiny_model, utterances):
    """rs_host_alloc: a batch laid out back to back in page-locked memory is copied to the device without the staging
    memcpy; out-of-order views of the same block fall back to staging.  Results are identical either way."""
    from rhasspy_speech_b200 import _lib
    dec = _lib.Decoder(_lib.Model(tiny_model.final_mdl, tiny_model.online_conf, 0), _lib.Graph(tiny_model.hclg, tiny_model.words_txt, 0))
    batch = list(utterances) + [np.zeros(0, np.int16), utterances[0][:399]]
    want = dec.decode_pcm(batch)
    pa = _lib.PinnedAudio.from_utterances(batch)
    got = dec.decode_pcm(pa.views)
    assert got.words == want.words and list(got.num_frames) == list(want.num_frames)
    h2d_direct = dec.timings()["h2d_bytes"]
    assert dec.decode_pcm(pa).words == want.words      # the block itself: cached argument arrays
    rev = dec.decode_pcm(pa.views[::-1])              # not in block order: staged like any other buffer
    assert rev.words == want.words[::-1] and dec.timings()["h2d_bytes"] == h2d_direct
    pa.close()


def test_device_pool_gives_the_single_device_results(tiny_model, utterances, synth, tmp_path):
    """SURVEY 8e on the product path: a transcriber over several engines (here two on device 0 -- and every visible device
    when the box has more) deals a request list longest-first, runs the shares concurrently and returns what one device
    returns; concurrent single requests and streams spread over the engines."""
    import rhasspy_speech_b200 as pkg
    from rhasspy_speech_b200 import _lib
    graph_dir = os.path.dirname(tiny_model.hclg)
    wavs = []
    for i in range(24):
        w = os.path.join(str(tmp_path), "p%03d.wav" % i)
        synth.write_wav(w, utterances[i % len(utterances)][:16000 + 997 * i])
        wavs.append(w)
    one = pkg.KaldiNnet3WavTranscriber(tiny_model.model_dir, graph_dir, None)
    want = asyncio.run(one.async_transcribe_many(wavs, tmp_path, nbest=2))
    devices = [0, 0] if _lib.device_count() < 2 else list(range(_lib.device_count()))
    pool = pkg.KaldiNnet3WavTranscriber(tiny_model.model_dir, graph_dir, None, device=devices)
    engines = pool._get_engines()
    assert len(engines) == len(devices) and len({id(e) for e in engines}) == len(devices)
    assert asyncio.run(pool.async_transcribe_many(wavs, tmp_path, nbest=2)) == want

    async def burst():
        return await asyncio.gather(*[pool.async_transcribe(w, tmp_path, nbest=2) for w in wavs])
    before = [len(e.batcher.batches) for e in engines]
    assert asyncio.run(burst()) == want
    assert sum(1 for e, b in zip(engines, before) if len(e.batcher.batches) > b) >= 2      # more than one engine took work
    # streams: sticky to the least-loaded engine
    sp = pkg.KaldiNnet3StreamTranscriber(tiny_model.model_dir, graph_dir, None, device=devices)

    async def chunks(pcm):
        raw = np.asarray(pcm, dtype="<i2").tobytes()
        for o in range(0, len(raw), 2560):
            yield raw[o:o + 2560]
            await asyncio.sleep(0)

    async def streams():
        return await asyncio.gather(*[sp.async_transcribe(chunks(utterances[i % len(utterances)]), tmp_path) for i in range(12)])
    s1 = pkg.KaldiNnet3StreamTranscriber(tiny_model.model_dir, graph_dir, None)

    async def lone():
        return [await s1.async_transcribe(chunks(utterances[i % len(utterances)]), tmp_path) for i in range(len(utterances))]
    got, ref = asyncio.run(streams()), asyncio.run(lone())
    assert got == [ref[i % len(utterances)] for i in range(12)]
    assert all(e.open_streams == 0 for e in sp._get_engines())
