import tempfile

import torch


def tiny_modernbert_dir(vocab_size=50000, hidden=128, layers=2, heads=4, seed=0, bias_std=0.02, bias_offset=0.0):
    """BASELINE config 1 trunk: random-init ModernBERT (2 layers, hidden 128, vocab 50k) saved to a temp dir so
    that SPLADEModernBERT(model_name=dir) loads it offline (SURVEY.md section 8(c) gotcha 2)."""
    from transformers import AutoModelForMaskedLM, ModernBertConfig
    cfg = ModernBertConfig(vocab_size=vocab_size, hidden_size=hidden, intermediate_size=hidden * 3 // 2,
                           num_hidden_layers=layers, num_attention_heads=heads, pad_token_id=vocab_size - 1,
                           bos_token_id=0, eos_token_id=1, cls_token_id=0, sep_token_id=1,
                           max_position_embeddings=512)
    torch.manual_seed(seed)
    hf = AutoModelForMaskedLM.from_config(cfg)
    with torch.no_grad():  # random-init decoder.bias is all zeros: exercise the bias path
        hf.decoder.bias.copy_(torch.randn(vocab_size, generator=torch.Generator().manual_seed(2)) * bias_std + bias_offset)
    d = tempfile.mkdtemp(prefix="tiny_modernbert_")
    hf.save_pretrained(d)
    return d
