#!/usr/bin/env python
"""Writes a random-init checkpoint the reference's generate.py can load offline:

    python tools/write_random_checkpoint.py /tmp/ckpt --model stories110M [--dir-name Llama-2-stories110M]

creates <out>/<org>/<dir-name>/model.pth (gpt-fast key names, bf16, normal(0, 0.02)) and, next to it, a tiny
SentencePiece `tokenizer.model` trained on the spot (no network): generate.py asserts that the checkpoint file exists
and loads the tokenizer from its directory (reference generate.py:68-76, 97-104), and takes the model shape from the
directory name (generation_utils.py:537).  A directory name containing "llama-2" makes the reference pick its
SentencePiece wrapper (tokenizer.py:254-259)."""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
TOK_VOCAB = 64
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out", type=Path)
    ap.add_argument("--model", default="stories110M")
    ap.add_argument("--dir-name", default=None)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    from cold_compress_b200.model import ModelArgs, Transformer
    name = a.dir_name or f"Llama-2-{a.model}"
    d = a.out / "random-init" / name
    d.mkdir(parents=True, exist_ok=True)
    torch.manual_seed(a.seed)
    model = Transformer(ModelArgs.from_name(a.model))
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "norm" in n:
                p.fill_(1.0)
            else:
                p.normal_(0.0, 0.02)
    sd = {k: v.to(torch.bfloat16) for k, v in model.state_dict().items()}
    # the tokenizer below knows TOK_VOCAB pieces: rows of the output head beyond it are zero, so greedy decoding
    # (arg-max over logits whose first TOK_VOCAB entries are random and the rest exactly 0) only emits ids that the
    # reference's tokenizer.decode (generate.py:136) accepts
    sd["output.weight"][TOK_VOCAB:] = 0
    sd["output.weight"][:TOK_VOCAB] *= 8
    sd["output.weight"][:3] = 0   # unk / bos / eos: an early eos would end generate.py's loop before max_new_tokens
    torch.save(sd, d / "model.pth")
    import sentencepiece as spm
    corpus = d / "corpus.txt"
    words = ["the", "cache", "keeps", "heavy", "hitters", "and", "recent", "tokens", "while", "old", "keys", "are", "evicted",
             "attention", "scores", "decide", "which", "slot", "is", "overwritten", "hello", "my", "name", "is", "a", "b200"]
    gen = torch.Generator().manual_seed(a.seed)
    lines = [" ".join(words[int(i)] for i in torch.randint(0, len(words), (12,), generator=gen)) for _ in range(400)]
    corpus.write_text("\n".join(lines) + "\n")
    spm.SentencePieceTrainer.train(input=str(corpus), model_prefix=str(d / "tokenizer"), vocab_size=TOK_VOCAB, model_type="bpe",
                                   bos_id=1, eos_id=2, unk_id=0, pad_id=-1, minloglevel=2)
    corpus.unlink()
    (d / "tokenizer.vocab").unlink(missing_ok=True)
    print(d / "model.pth")


if __name__ == "__main__":
    main()
