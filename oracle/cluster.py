"""Speaker assignment restated in numpy — test infrastructure only.

* EmbeddingManager: pyannote-rs' online leader clustering exactly as the crate drives it (reference src/transcribe.rs:342,
  480-492): strict `>` threshold, ids from 1, stored embedding never updated, cap -> best match regardless of threshold.
  (pyannote-rs iterates a HashMap, i.e. ties between speakers are order-dependent upstream; ties resolve toward the lowest id.)
* leader_labels(S, ...): the same policy as a pure function of the pairwise similarity matrix.
* agglomerative_labels(S, thr): average-linkage agglomerative clustering (north-star), merges while max similarity > thr;
  cross-checked against scipy.cluster.hierarchy in tests/test_oracle_cluster.py."""
import numpy as np

F = np.float32


def cosine_similarity(a, b):
    a = np.asarray(a, F)
    b = np.asarray(b, F)
    dot = F(0)
    na = F(0)
    nb = F(0)
    for x, y in zip(a, b):  # sequential fp32 accumulation, as ndarray's dot on small vectors would not guarantee
        dot = F(dot + F(x * y))
        na = F(na + F(x * x))
        nb = F(nb + F(y * y))
    na, nb = F(np.sqrt(na)), F(np.sqrt(nb))
    if na == 0 or nb == 0:
        return F(0)
    return F(dot / F(na * nb))


def cosine_matrix(E):
    E = np.asarray(E, np.float64)
    n = np.sqrt((E * E).sum(1))
    S = (E @ E.T) / np.maximum(np.outer(n, n), 1e-300)
    S[np.outer(n, n) == 0] = 0.0
    return S.astype(F)


class EmbeddingManager:
    def __init__(self, max_speakers):
        self.max_speakers = max_speakers
        self.speakers = {}
        self.next_id = 1

    def search_speaker(self, emb, threshold):
        best_id, best = None, F(threshold)
        for sid in sorted(self.speakers):
            s = cosine_similarity(emb, self.speakers[sid])
            if s > best:
                best_id, best = sid, s
        if best_id is not None:
            return best_id
        if len(self.speakers) < self.max_speakers:
            sid = self.next_id
            self.next_id += 1
            self.speakers[sid] = np.asarray(emb, F).copy()
            return sid
        return None

    def get_best_speaker_match(self, emb):
        if not self.speakers:
            raise ValueError("no speakers")
        best_id, best = 0, -np.inf
        for sid in sorted(self.speakers):
            s = cosine_similarity(emb, self.speakers[sid])
            if s > best:
                best_id, best = sid, s
        return best_id

    def assign(self, emb, threshold):
        """The crate's policy (src/transcribe.rs:482-492). Returns id or None ('?')."""
        if len(self.speakers) == self.max_speakers:
            return self.get_best_speaker_match(emb)
        return self.search_speaker(emb, threshold)


def leader_labels(S, threshold, max_speakers):
    S = np.asarray(S, F)
    N = S.shape[0]
    rep, labels = [], np.zeros(N, np.int32)
    for i in range(N):
        best_id = 0
        if len(rep) == max_speakers:
            best = -np.inf
            for k, r in enumerate(rep):
                if S[i, r] > best:
                    best, best_id = S[i, r], k + 1
        else:
            best = F(threshold)
            for k, r in enumerate(rep):
                if S[i, r] > best:
                    best, best_id = S[i, r], k + 1
            if not best_id and len(rep) < max_speakers:
                rep.append(i)
                best_id = len(rep)
        labels[i] = best_id
    return labels


def agglomerative_labels(S, threshold):
    W = np.array(S, F, copy=True)
    N = W.shape[0]
    cnt = np.ones(N, np.int64)
    parent = np.arange(N)
    iu = np.triu(np.ones((N, N), bool), 1)
    for _ in range(N - 1):
        alive = cnt > 0
        mask = iu & alive[:, None] & alive[None, :]
        if not mask.any():
            break
        Wm = np.where(mask, W, -np.inf)
        f = int(np.argmax(Wm))  # first maximum in row-major order
        ci, cj = divmod(f, N)
        if not (W[ci, cj] > F(threshold)):
            break
        ni, nj = F(cnt[ci]), F(cnt[cj])
        den = F(ni + nj)
        for k in range(N):
            if k == ci or k == cj or cnt[k] == 0:
                continue
            a = W[min(ci, k), max(ci, k)]
            c = W[min(cj, k), max(cj, k)]
            W[min(ci, k), max(ci, k)] = F(F(F(ni * a) + F(nj * c)) / den)
        cnt[ci] += cnt[cj]
        cnt[cj] = 0
        parent[parent == cj] = ci
    labels = np.zeros(N, np.int32)
    nxt = 1
    for i in range(N):
        if parent[i] == i:
            labels[i] = nxt
            nxt += 1
        else:
            labels[i] = labels[parent[i]]
    return labels
