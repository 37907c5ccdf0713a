"""Integer id helpers with the reference's names and semantics (genie/factorization_utils.py:55-106).
Pure index arithmetic on whatever device the ids live on (host-side glue, not part of the kernel path;
inside the kernels the same arithmetic is `id % V`, `id / V`)."""
import torch

from .config import nth_root  # noqa: F401


def factorize_token_ids(token_ids: torch.LongTensor, num_factored_vocabs: int = 2,
                        factored_vocab_size: int = 512) -> torch.LongTensor:
    powers = factored_vocab_size ** torch.arange(num_factored_vocabs, device=token_ids.device)
    return (token_ids.unsqueeze(-1) // powers) % factored_vocab_size


def unfactorize_token_ids(factored_token_ids: torch.LongTensor, num_factored_vocabs: int = 2,
                          factored_vocab_size: int = 512) -> torch.LongTensor:
    powers = factored_vocab_size ** torch.arange(num_factored_vocabs, device=factored_token_ids.device)
    return (factored_token_ids * powers).sum(dim=-1)


def factorize_labels(labels_THW: torch.LongTensor, num_factored_vocabs: int = 2,
                     factored_vocab_size: int = 512) -> torch.LongTensor:
    f = factorize_token_ids(labels_THW, num_factored_vocabs, factored_vocab_size)
    return f.permute(0, 4, 1, 2, 3)
