"""CLIP's byte-level BPE tokeniser, native (host-side string processing; no transformers / open_clip import on the serving path).

Both text towers of the reference tokenise with the SAME vocabulary and algorithm — OpenAI CLIP's `simple_tokenizer` — reached
through two packages: `transformers.CLIPTokenizer` for CLIP-L (sgm/modules/encoders/modules.py:462, 485-494: truncation to 77,
padded with <|endoftext|>) and `open_clip.tokenize` for bigG (modules.py:554: [SOT] ids [EOT], truncation keeps EOT last, ZERO
padded). The algorithm restated here: clean the text (html-unescape twice, NFC, collapse whitespace, lower-case — ftfy's
`fix_text` too when that package is importable, as the reference's environment has it), split with CLIP's pattern, map each
piece's UTF-8 bytes to the printable byte alphabet, merge pairs greedily by merge rank with `</w>` closing every piece, look
the sub-words up in the vocabulary.

Vocabulary files (none exist offline — tests build a synthetic one): a Hugging Face directory with `vocab.json` + `merges.txt`
(what CKPT_PTH.SDXL_CLIP1_PATH points at) or open_clip's `bpe_simple_vocab_16e6.txt.gz`. Pinned against the independent Rust
implementation behind transformers' CLIPTokenizer on a trained synthetic vocabulary (tests/test_clip_bpe.py).
"""
import gzip
import html
import json
import os
import unicodedata
from functools import lru_cache

try:
    import regex as re          # \p{L} / \p{N} classes: the third-party `regex` module, like OpenAI's simple_tokenizer
except ImportError as _e:        # only tokenising STRINGS needs it; token-id inputs do not
    raise ImportError("supir_b200.clip_bpe needs the `regex` package (unicode property classes in CLIP's split pattern)") from _e

SOT, EOT = "<|startoftext|>", "<|endoftext|>"
PATTERN = re.compile(r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""", re.IGNORECASE)


@lru_cache()
def bytes_to_unicode():
    """The reversible byte -> printable-character table of GPT-2 / CLIP byte-level BPE."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("\xa1"), ord("\xac") + 1)) + list(range(ord("\xae"), ord("\xff") + 1))
    cs = bs[:]
    n = 0
    for b in range(2 ** 8):
        if b not in bs:
            bs.append(b)
            cs.append(2 ** 8 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


def clean(text):
    try:
        import ftfy
        text = ftfy.fix_text(text)
    except ImportError:
        pass
    text = html.unescape(html.unescape(text))
    text = unicodedata.normalize("NFC", text)
    return re.sub(r"\s+", " ", text).strip().lower()


class ClipBPE:
    def __init__(self, vocab, merges):
        """vocab: {token string: id}; merges: [(left, right), ...] in rank order."""
        self.encoder = dict(vocab)
        self.ranks = {tuple(m): i for i, m in enumerate(merges)}
        self.byte_encoder = bytes_to_unicode()
        self.sot, self.eot = self.encoder[SOT], self.encoder[EOT]
        self.cache = {SOT: SOT, EOT: EOT}

    # ---- loading ----
    @classmethod
    def from_path(cls, path):
        """`path`: a directory holding vocab.json + merges.txt (Hugging Face layout) or a bpe_simple_vocab_16e6.txt.gz file."""
        if os.path.isdir(path):
            with open(os.path.join(path, "vocab.json"), encoding="utf-8") as f:
                vocab = json.load(f)
            with open(os.path.join(path, "merges.txt"), encoding="utf-8") as f:
                lines = [ln for ln in f.read().split("\n") if ln and not ln.startswith("#version")]
            return cls(vocab, [tuple(ln.split()) for ln in lines])
        with gzip.open(path) as f:                                      # open_clip / OpenAI simple_tokenizer.py layout
            lines = f.read().decode("utf-8").split("\n")
        merges = [tuple(m.split()) for m in lines[1:49152 - 256 - 2 + 1]]
        alphabet = list(bytes_to_unicode().values())
        tokens = alphabet + [c + "</w>" for c in alphabet] + ["".join(m) for m in merges] + [SOT, EOT]
        return cls({t: i for i, t in enumerate(tokens)}, merges)

    @staticmethod
    def available(path):
        return isinstance(path, str) and ((os.path.isdir(path) and os.path.isfile(os.path.join(path, "vocab.json"))
                                           and os.path.isfile(os.path.join(path, "merges.txt")))
                                          or (os.path.isfile(path) and path.endswith(".txt.gz")))

    # ---- the algorithm ----
    def bpe(self, token):
        if token in self.cache:
            return self.cache[token]
        word = tuple(token[:-1]) + (token[-1] + "</w>",)
        while len(word) > 1:
            pairs = set(zip(word, word[1:]))
            best = min(pairs, key=lambda p: self.ranks.get(p, float("inf")))
            if best not in self.ranks:
                break
            first, second = best
            out, i = [], 0
            while i < len(word):
                if i < len(word) - 1 and word[i] == first and word[i + 1] == second:
                    out.append(first + second)
                    i += 2
                else:
                    out.append(word[i])
                    i += 1
            word = tuple(out)
        res = " ".join(word)
        self.cache[token] = res
        return res

    def encode(self, text):
        ids = []
        unk = self.eot                                                   # CLIPTokenizer's unk_token is <|endoftext|>
        for piece in PATTERN.findall(clean(text)):
            piece = "".join(self.byte_encoder[b] for b in piece.encode("utf-8"))
            ids.extend(self.encoder.get(t, unk) for t in self.bpe(piece).split(" "))
        return ids

    # ---- the two layouts the reference feeds its towers ----
    def tokenize_hf(self, texts, max_length=77):
        """CLIPTokenizer(text, truncation=True, max_length=77, padding='max_length'): [SOT] ids[:75] [EOT], padded with EOT."""
        rows = []
        for t in texts:
            ids = [self.sot] + self.encode(t)[:max_length - 2] + [self.eot]
            rows.append(ids + [self.eot] * (max_length - len(ids)))
        return rows

    def tokenize_open_clip(self, texts, context_length=77):
        """open_clip.tokenize: [SOT] ids [EOT]; longer rows are cut to the context with EOT as their last token; zero padded."""
        rows = []
        for t in texts:
            ids = [self.sot] + self.encode(t) + [self.eot]
            if len(ids) > context_length:
                ids = ids[:context_length]
                ids[-1] = self.eot
            rows.append(ids + [0] * (context_length - len(ids)))
        return rows
