"""supir_b200/clip_bpe.py (native CLIP byte-level BPE) against the independent Rust implementation behind
transformers.CLIPTokenizer, on a SYNTHETIC vocabulary trained here with the `tokenizers` library on this repository's own
documents (no real CLIP vocabulary exists offline): ids of every test string, the CLIP-L padding / truncation layout, and
open_clip's layout rules. Also the loader for open_clip's bpe_simple_vocab_16e6.txt.gz layout."""
import gzip
import os
import random

import pytest
import torch

tokenizers = pytest.importorskip("tokenizers")
transformers = pytest.importorskip("transformers")
from supir_b200 import clip_bpe  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATTERN = r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+"""


def corpus():
    lines = []
    for name in ("SURVEY.md", "DESIGN.md", "README.md", "INTEGRATION.md", "BASELINE.md"):
        with open(os.path.join(ROOT, name), encoding="utf-8") as f:
            lines += [ln for ln in f.read().split("\n") if ln.strip()]
    return lines


@pytest.fixture(scope="module")
def trained(tmp_path_factory):
    from tokenizers import Regex, Tokenizer, models, normalizers, pre_tokenizers, trainers
    tok = Tokenizer(models.BPE(end_of_word_suffix="</w>", continuing_subword_prefix="", unk_token="<|endoftext|>"))
    tok.normalizer = normalizers.Sequence([normalizers.NFC(), normalizers.Replace(Regex(r"\s+"), " "), normalizers.Lowercase()])
    tok.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Split(Regex(PATTERN), behavior="removed", invert=True),
                                                 pre_tokenizers.ByteLevel(add_prefix_space=False)])
    trainer = trainers.BpeTrainer(vocab_size=4000, special_tokens=[], initial_alphabet=pre_tokenizers.ByteLevel.alphabet(),
                                  end_of_word_suffix="</w>", show_progress=False)
    tok.train_from_iterator(corpus(), trainer)
    d = tmp_path_factory.mktemp("bpe")
    tok.model.save(str(d))                                                 # vocab.json + merges.txt (the Hugging Face layout)
    import json
    with open(d / "vocab.json", encoding="utf-8") as f:
        vocab = json.load(f)
    for sp in (clip_bpe.SOT, clip_bpe.EOT):                                # CLIP keeps its two specials at the END of the table
        vocab[sp] = len(vocab)
    with open(d / "vocab.json", "w", encoding="utf-8") as f:
        json.dump(vocab, f, ensure_ascii=False)
    with open(d / "merges.txt", encoding="utf-8") as f:
        merges = [tuple(ln.split()) for ln in f.read().split("\n") if ln and not ln.startswith("#version")]
    hf = transformers.CLIPTokenizer(vocab=vocab, merges=[tuple(m) for m in merges])
    return str(d), vocab, merges, hf


STRINGS = ["a photo of a cat", "Cinematic, High Contrast, highly detailed, taken using a Canon EOS R camera, hyper detailed photo - realistic maximum detail",
           "it's the dog's 2nd birthday — don't   you've  they'll I'm we'd", "TILE_SIZE=128,stride=64; 4096x4096 -> 16.78 MP!!! (bf16/fp16)",
           "  leading and trailing   whitespace \t\n newline ", "naïve café déjà-vu Ångström", "日本語のテキスト と emoji 🙂🚀",
           "", "<|startoftext|> inside <|endoftext|> text", "ＦＵＬＬ　ＷＩＤＴＨ １２３", "tcgen05.mma cta_group::2 UTCHMMA.2CTA 1e-5 0.13025"]


def test_ids_match_the_tokenizers_backend(trained):
    path, vocab, merges, hf = trained
    mine = clip_bpe.ClipBPE.from_path(path)
    assert clip_bpe.ClipBPE.available(path) and not clip_bpe.ClipBPE.available(os.path.join(path, "nope"))
    rng = random.Random(0)
    words = [w for ln in corpus()[:400] for w in ln.split()]
    extra = [" ".join(rng.choice(words) for _ in range(rng.randint(1, 120))) for _ in range(60)]
    for s in STRINGS + extra:
        want = hf(s, add_special_tokens=False)["input_ids"]
        assert mine.encode(s) == list(want), s
    # html entities: OpenAI's / open_clip's cleaner (and ftfy inside the reference's transformers 4.28 tokeniser) unescape them,
    # the tokenizers backend of this image's transformers does not — follow the reference's environment
    assert mine.encode("&amp;lt;b&amp;gt; html &quot;entities&quot;") == mine.encode('<b> html "entities"')
    texts = STRINGS[:6] + extra[:10]
    want = hf(texts, truncation=True, max_length=77, return_length=True, return_overflowing_tokens=False, padding="max_length",
              return_tensors="pt")["input_ids"]
    got = torch.tensor(mine.tokenize_hf(texts, 77))
    assert torch.equal(got, want)                                         # [SOT] ids [EOT], EOT padding, truncation to 77


def test_open_clip_layout_and_gz_loader(trained, tmp_path):
    path, vocab, merges, hf = trained
    mine = clip_bpe.ClipBPE.from_path(path)
    long_text = "word " * 200
    rows = mine.tokenize_open_clip(["a cat", long_text, ""], 77)
    ids = mine.encode("a cat")
    assert rows[0] == [mine.sot] + ids + [mine.eot] + [0] * (77 - len(ids) - 2)
    assert len(rows[1]) == 77 and rows[1][0] == mine.sot and rows[1][-1] == mine.eot and 0 not in rows[1]
    assert rows[2] == [mine.sot, mine.eot] + [0] * 75
    # open_clip's vocabulary file: a header line, then one merge per line; the token table is rebuilt from the byte alphabet.
    # simple_tokenizer reads exactly 49152 - 256 - 2 merges; write that many (real merges first, then never-matching filler).
    n = 49152 - 256 - 2
    filler = [(f"Ā{i}", f"ā{i}") for i in range(n - len(merges))]
    gz = tmp_path / "bpe_simple_vocab_16e6.txt.gz"
    with gzip.open(gz, "wb") as f:
        f.write(("#version: synthetic\n" + "\n".join(" ".join(m) for m in list(merges) + filler) + "\n").encode("utf-8"))
    oc = clip_bpe.ClipBPE.from_path(str(gz))
    assert oc.sot == 49406 and oc.eot == 49407 and len(oc.encoder) == 49408
    alphabet = list(clip_bpe.bytes_to_unicode().values())
    assert oc.encoder[alphabet[0]] == 0 and oc.encoder[alphabet[0] + "</w>"] == 256 and oc.encoder["".join(merges[0])] == 512
    inv = {v: k for k, v in oc.encoder.items()}
    for s in STRINGS[:4]:                                                  # same merges -> same segmentation; its own (complete) id table
        pieces = ["".join(oc.byte_encoder[b] for b in p.encode("utf-8")) for p in clip_bpe.PATTERN.findall(clip_bpe.clean(s))]
        want = [t for p in pieces for t in mine.bpe(p).split(" ")]
        assert [inv[i] for i in oc.encode(s)] == want


def test_embedders_tokenise_natively_when_the_vocabulary_is_on_disk(trained):
    from supir_b200 import conditioner as C
    from weights import COND_G, COND_L
    path, vocab, merges, hf = trained
    l = C.FrozenCLIPEmbedder(layer="hidden", layer_idx=1, arch=dict(COND_L, layers=2))
    g = C.FrozenOpenCLIPEmbedder2(arch="ViT-bigG-14", layer="penultimate", legacy=False, always_return_pooled=True, text_cfg=dict(COND_G, layers=2),
                                  tokenizer_path=path)
    l.tokenizer_path = path
    texts = ["a photo of a cat", "word " * 100]
    bpe = clip_bpe.ClipBPE.from_path(path)
    assert torch.equal(l.tokenize(texts), torch.tensor(bpe.tokenize_hf(texts, 77)))
    assert torch.equal(g.tokenize(texts), torch.tensor(bpe.tokenize_open_clip(texts, 77)))


def test_fuzz_against_the_tokenizers_backend(trained):
    """800 random strings over ASCII punctuation, accented Latin, CJK, emoji (with modifiers), all kinds of whitespace, typographic
    quotes, non-Latin digits, title-case ligatures and combining marks: identical ids ('&' left out: html unescaping differs by
    design, see above)."""
    path, vocab, merges, hf = trained
    mine = clip_bpe.ClipBPE.from_path(path)
    rng = random.Random(1)
    pools = ["".join(chr(c) for c in range(32, 127) if chr(c) != "&"), "äöüßéèêñçøåÆ¿¡", "日本語中文한국어", "🙂🚀👍🏽", "\t\n\r\x0b\x0c  ", "’‘“”–—…",
             "١٢٣४५६", "ǅǈﬁﬀ", "́̈"]
    for _ in range(800):
        s = "".join(rng.choice(rng.choice(pools)) for _ in range(rng.randint(0, 40)))
        assert mine.encode(s) == list(hf(s, add_special_tokens=False)["input_ids"]), repr(s)
