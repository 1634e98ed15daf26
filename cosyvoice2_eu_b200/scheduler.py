"""Event-driven chunk scheduler for many concurrent streaming sessions.

The reference runs one polling loop per request (`CosyVoice2Model.tts`, cosyvoice/cli/model.py:351-381): every 0.1 s it looks
at a Python list the LLM thread appends to under a lock, and when `token_hop_len + pre_lookahead_len` new tokens are there it
calls `token2wav` for that one session.  Here the producers (`push`) wake the consumer through a condition variable -- no poll
quantum -- and one `step()` takes EVERY session that has a chunk ready and runs them as one ragged batch through
`B200Token2Wav.token2wav_stream_batch` (one flow launch sequence and one vocoder launch sequence for all of them).

The chunk arithmetic is the reference's, per session:
  * the first hop is `token_hop_len + prompt_token_pad`, `prompt_token_pad = ceil(P / hop) * hop - P` (model.py:353, 356);
  * a non-final chunk is dispatched when `len(tokens) - token_offset >= this_hop + pre_lookahead_len` and sees
    `tokens[:token_offset + this_hop + pre_lookahead_len]` (model.py:357-358), then `token_offset += this_hop` (:367);
  * once the producer has finished and fewer than that remain, ONE final call sees all tokens with `finalize=True`
    (model.py:369-380; it runs with full attention, the quirk SURVEY.md section 8a lists), and the session's cache is dropped
    (model.py:395-396).

  * the loop's break test (model.py:369) reuses the hop computed before `token_offset` moved.  That only differs from the
    recomputed hop right after the FIRST chunk (hop + pad), and only if the producer had already finished: the reference then
    finalizes as soon as fewer than hop + pad + lookahead tokens remain.  Reproduced here (`_Session.force_final`), and pinned
    against the reference's own loop in tests/golden/tts_schedule.npz.

Host logic only: no arithmetic happens here, and nothing here needs a GPU (tests/test_host_logic.py drives it with a stub).
"""
import math
import threading


def chunk_schedule(n_tokens, n_prompt, hop=25, lookahead=3, all_tokens_ready=False):
    """The (n_tokens_visible, token_offset, finalize) calls CosyVoice2Model.tts(stream=True) makes for one session
    (model.py:351-381).  `all_tokens_ready`: the producer had finished before the first chunk went out (vc_job, model.py:141-143)."""
    pad = int(math.ceil(n_prompt / hop) * hop - n_prompt)
    calls, off = [], 0
    while True:
        this_hop = hop + pad if off == 0 else hop
        if n_tokens - off < this_hop + lookahead:
            break
        calls.append((off + this_hop + lookahead, off, False))
        off += this_hop
        if all_tokens_ready and n_tokens - off < this_hop + lookahead:
            break
    calls.append((n_tokens, off, True))
    return calls


class _Session:
    __slots__ = ("uuid", "prompt_token", "prompt_feat", "embedding", "tokens", "token_offset", "pad", "ended", "done", "force_final")

    def __init__(self, uuid, prompt_token, prompt_feat, embedding, hop):
        self.uuid = uuid
        self.prompt_token, self.prompt_feat, self.embedding = prompt_token, prompt_feat, embedding
        self.tokens = []
        self.token_offset = 0
        n_prompt = int(prompt_token.shape[1])
        self.pad = int(math.ceil(n_prompt / hop) * hop - n_prompt)
        self.ended = False
        self.done = False
        self.force_final = False


class StreamScheduler:
    """open() a session, push() speech tokens as the LLM produces them, close() when it has finished; step() (or run()) turns
    whatever is ready into audio chunks: a list of (uuid, speech [1, L], is_final)."""

    def __init__(self, t2w, token_hop_len=25, pre_lookahead_len=None, group=None):
        """group: optional StreamGroup (B200Flow.open_stream_group) -- non-final chunks then run incrementally on the device state
        it holds (row F1) instead of recomputing the prefix."""
        self.t2w = t2w
        self.group = group
        self.token_hop_len = int(token_hop_len)
        self.pre_lookahead_len = int(t2w.flow.pre_lookahead_len if pre_lookahead_len is None else pre_lookahead_len)
        self.sessions = {}
        self.cv = threading.Condition()

    # ---------------------------------------------------------------- producer side (LLM threads)
    def open(self, uuid, prompt_token, prompt_feat, embedding):
        with self.cv:
            if uuid in self.sessions:
                raise KeyError(f"session {uuid!r} is already open")
            self.sessions[uuid] = _Session(uuid, prompt_token, prompt_feat, embedding, self.token_hop_len)
            self.t2w.hift_cache_dict[uuid] = None                    # model.py:345

    def push(self, uuid, tokens):
        """Append speech tokens (an int or an iterable of ints), like the LLM job does (model.py:122-139)."""
        with self.cv:
            s = self.sessions[uuid]
            if s.ended:
                raise RuntimeError(f"session {uuid!r} was closed")
            if isinstance(tokens, int):
                s.tokens.append(tokens)
            else:
                s.tokens.extend(int(t) for t in tokens)
            if self._ready(s):
                self.cv.notify_all()

    def close(self, uuid):
        """The producer has finished (llm_end_dict[uuid] = True, model.py:139)."""
        with self.cv:
            self.sessions[uuid].ended = True
            self.cv.notify_all()

    # ---------------------------------------------------------------- chunk arithmetic (model.py:353-369)
    def _this_hop(self, s):
        return self.token_hop_len + s.pad if s.token_offset == 0 else self.token_hop_len

    def _ready(self, s):
        return not s.force_final and len(s.tokens) - s.token_offset >= self._this_hop(s) + self.pre_lookahead_len

    def _final_due(self, s):
        return s.force_final or (s.ended and not self._ready(s))

    def pending(self):
        with self.cv:
            return any(self._ready(s) or self._final_due(s) for s in self.sessions.values())

    def wait(self, timeout=None):
        """Block until some session has a chunk (or its final call) due; returns False on timeout."""
        with self.cv:
            return self.cv.wait_for(lambda: any(self._ready(s) or self._final_due(s) for s in self.sessions.values()), timeout)

    # ---------------------------------------------------------------- consumer side
    def step(self, noises=None):
        """One batched step: every session with a non-final chunk ready goes into one token2wav_stream_batch(finalize=False)
        call, every finished session into one finalize=True call.  `noises` (parity runs only): {uuid: NSF noise tensor}."""
        import torch
        with self.cv:
            chunk, final = [], []
            for s in self.sessions.values():
                if self._ready(s):
                    n_vis = s.token_offset + self._this_hop(s) + self.pre_lookahead_len
                    chunk.append((s, n_vis, s.token_offset))
                    hop = self._this_hop(s)
                    s.token_offset += hop
                    # model.py:369: the break test still holds the hop of the chunk just emitted
                    if s.ended and len(s.tokens) - s.token_offset < hop + self.pre_lookahead_len:
                        s.force_final = True
                elif self._final_due(s):
                    final.append((s, len(s.tokens), s.token_offset))
                    s.done = True
            snap = [(grp, fin, [(s, torch.tensor(s.tokens[:n_vis], dtype=torch.int32).unsqueeze(0), off) for s, n_vis, off in grp])
                    for grp, fin in ((chunk, False), (final, True)) if grp]
        out = []
        for _, fin, items in snap:
            reqs = [dict(token=tok, prompt_token=s.prompt_token, prompt_feat=s.prompt_feat, embedding=s.embedding, token_offset=off,
                         uuid=s.uuid) for s, tok, off in items]
            nz = None if noises is None else [noises[s.uuid] for s, _, _ in items]
            kw = {} if self.group is None else {"group": self.group}
            speeches = self.t2w.token2wav_stream_batch(reqs, finalize=fin, noises=nz, **kw)
            out.extend((s.uuid, sp, fin) for (s, _, _), sp in zip(items, speeches))
        with self.cv:
            for uuid in [u for u, s in self.sessions.items() if s.done]:
                del self.sessions[uuid]
                self.t2w.hift_cache_dict.pop(uuid, None)             # model.py:395-396
        return out

    def run(self, on_chunk, idle_timeout=None):
        """Serve until every open session has delivered its final chunk: on_chunk(uuid, speech, is_final) per chunk."""
        while True:
            with self.cv:
                if not self.sessions:
                    return
            if not self.wait(idle_timeout):
                return
            for uuid, speech, fin in self.step():
                on_chunk(uuid, speech, fin)
