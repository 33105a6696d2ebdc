"""Word-level language-model data: dictionary, corpus, batchify, get_batch.

Capabilities of reference ``dataloader.py:120-172`` and ``utils.py:7-11``: a vocabulary built
over train+valid+test in file order with ``<eos>`` appended to every line (33 278 types on
wikitext-2), column-major ``batchify`` and bptt slicing.  Differences: the corpus is tokenised
ONCE per run and cached as a ``.pt`` next to the text (the reference re-tokenises every epoch,
SURVEY D11), ``get_batch`` honours ``bptt`` (D16), and when no corpus is on disk a synthetic
Zipfian stream with the same vocabulary size and split lengths is generated so shapes (and
therefore kernels and timings) are unchanged.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch

WIKITEXT2_VOCAB = 33278
WIKITEXT2_TOKENS = {"train": 2088628, "valid": 217646, "test": 245569}
_SEARCH = ("rnn_data/wikitext-2", "./data/wikitext-2", "/root/reference/rnn_data/wikitext-2")


class Dictionary:
    def __init__(self):
        self.word2idx: Dict[str, int] = {}
        self.idx2word: List[str] = []

    def add_word(self, word: str) -> int:
        i = self.word2idx.get(word)
        if i is None:
            i = len(self.idx2word)
            self.idx2word.append(word)
            self.word2idx[word] = i
        return i

    def __len__(self) -> int:
        return len(self.idx2word)


class Corpus:
    """``Corpus(path)`` tokenises ``train.txt / valid.txt / test.txt`` under ``path``."""

    def __init__(self, path: str):
        self.dictionary = Dictionary()
        self.train = self.tokenize(os.path.join(path, "train.txt"))
        self.valid = self.tokenize(os.path.join(path, "valid.txt"))
        self.test = self.tokenize(os.path.join(path, "test.txt"))

    def tokenize(self, path: str) -> torch.Tensor:
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        ids: List[int] = []
        add = self.dictionary.add_word
        with open(path, "r", encoding="utf8") as f:
            for line in f:
                for w in line.split():
                    ids.append(add(w))
                ids.append(add("<eos>"))
        return torch.tensor(ids, dtype=torch.int64)

    @property
    def ntokens(self) -> int:
        return len(self.dictionary)


class SyntheticCorpus:
    """Zipf-distributed token stream with wikitext-2's vocabulary size and split lengths."""

    def __init__(self, vocab: int = WIKITEXT2_VOCAB, sizes: Optional[Dict[str, int]] = None, seed: int = 1234):
        sizes = dict(WIKITEXT2_TOKENS if sizes is None else sizes)
        g = torch.Generator().manual_seed(seed)
        ranks = torch.arange(1, vocab + 1, dtype=torch.float64)
        prob = 1.0 / ranks
        prob /= prob.sum()
        self._vocab = vocab

        def draw(n):
            return torch.multinomial(prob, n, replacement=True, generator=g).to(torch.int64)
        self.train = draw(sizes["train"])
        self.valid = draw(sizes["valid"])
        self.test = draw(sizes["test"])
        # make sure the largest id is present so ntokens is exact
        self.train[0] = vocab - 1

    @property
    def ntokens(self) -> int:
        return self._vocab


def find_corpus_dir(root: str = "") -> Optional[str]:
    for cand in ((root,) if root else ()) + _SEARCH:
        if cand and os.path.isfile(os.path.join(cand, "train.txt")):
            return cand
    return None


_CACHE: Dict[str, object] = {}


def load_corpus(root: str = "", synthetic: Optional[bool] = None, sizes: Optional[Dict[str, int]] = None,
                seed: int = 1234):
    """Tokenise once per process (and cache on disk when writable)."""
    path = None if synthetic else find_corpus_dir(root)
    if path is None:
        if synthetic is False:
            raise FileNotFoundError("wikitext-2 not found; pass --corpus_root or --synthetic true")
        key = f"synthetic:{seed}:{sorted((sizes or {}).items())}"
        if key not in _CACHE:
            _CACHE[key] = SyntheticCorpus(sizes=sizes, seed=seed)
        return _CACHE[key]
    if path not in _CACHE:
        # the token cache lives under the user's cache directory, keyed by the corpus path -- never next to the corpus itself
        # (round 1 wrote `.dlb_tokens.pt` into /root/reference/rnn_data, a side effect on a tree this repo must not touch)
        import hashlib
        cache_dir = os.environ.get("DLB_CACHE_DIR") or os.path.join(os.path.expanduser("~"), ".cache", "dlb_b200")
        cache_file = os.path.join(cache_dir, "tokens_" + hashlib.sha1(os.path.realpath(path).encode()).hexdigest()[:16] + ".pt")
        corpus = None
        if os.path.isfile(cache_file):
            try:
                blob = torch.load(cache_file)
                corpus = Corpus.__new__(Corpus)
                corpus.dictionary = Dictionary()
                corpus.dictionary.idx2word = blob["idx2word"]
                corpus.dictionary.word2idx = {w: i for i, w in enumerate(blob["idx2word"])}
                corpus.train, corpus.valid, corpus.test = blob["train"], blob["valid"], blob["test"]
            except Exception:
                corpus = None
        if corpus is None:
            corpus = Corpus(path)
            try:
                os.makedirs(cache_dir, exist_ok=True)
                tmp = f"{cache_file}.{os.getpid()}.tmp"
                torch.save({"idx2word": corpus.dictionary.idx2word, "train": corpus.train,
                            "valid": corpus.valid, "test": corpus.test}, tmp)
                os.replace(tmp, cache_file)          # atomic: several ranks may tokenise at the same time
            except OSError:
                pass                         # unwritable cache directory: just re-tokenise next time
        _CACHE[path] = corpus
    return _CACHE[path]


def batchify(data: torch.Tensor, bsz: int) -> torch.Tensor:
    """[T] → [T//bsz, bsz], column *j* is a contiguous piece of the stream
    (reference ``dataloader.py:164-172``)."""
    bsz = int(bsz)
    nbatch = data.size(0) // bsz
    data = data.narrow(0, 0, nbatch * bsz)
    return data.view(bsz, -1).t().contiguous()


def get_batch(source: torch.Tensor, i: int, bptt: int = 35) -> Tuple[torch.Tensor, torch.Tensor]:
    """Rows ``i … i+bptt`` as input and the same rows shifted by one, flattened, as target
    (reference ``utils.py:7-11``, but honouring ``bptt``)."""
    seq_len = min(bptt, len(source) - 1 - i)
    data = source[i:i + seq_len]
    target = source[i + 1:i + 1 + seq_len].reshape(-1)
    return data, target
