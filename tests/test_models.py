"""Every model of the registry builds, runs a forward/backward at a reduced resolution and has the parameter count of
the architecture it names (the reference benchmarks torchvision's resnet/vgg/densenet, its own inceptionv4 and
transformers' BERT: dear/imagenet_benchmark.py:70-82, dear/bert_benchmark.py:60-75)."""
import pytest
import torch

from dear_pytorch_b200.models.registry import available, create, input_size

# torchvision / transformers parameter counts of the named architectures
EXPECTED = {"resnet18": 11_689_512, "resnet34": 21_797_672, "resnet50": 25_557_032, "resnet101": 44_549_160,
            "resnet152": 60_192_808, "vgg11": 132_863_336, "vgg16": 138_357_544, "vgg19": 143_667_240,
            "densenet121": 7_978_856, "densenet169": 14_149_480, "densenet201": 20_013_928}


def test_registry_lists_the_reference_models():
    names = set(available())
    assert {"resnet50", "vgg16", "densenet121", "inceptionv4", "bert", "bert_base", "mnist"} <= names
    assert input_size("inceptionv4") == 299 and input_size("resnet50") == 224


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_parameter_counts_match_the_named_architecture(name):
    with torch.device("meta"):
        model = create(name)
    assert sum(p.numel() for p in model.parameters()) == EXPECTED[name]


@pytest.mark.parametrize("name,size", [("resnet18", 64), ("vgg11", 32), ("densenet121", 64), ("inceptionv4", 96), ("mnist", 28)])
def test_forward_backward_at_reduced_resolution(name, size):
    torch.manual_seed(0)
    model = create(name)
    ch = 1 if name == "mnist" else 3
    x = torch.randn(2, ch, size, size)
    if name.startswith("vgg"):                         # the classifier expects 7x7 features: keep the spec'd input
        x = torch.randn(1, 3, 224, 224)
    out = model(x)
    assert out.shape[0] == x.shape[0] and torch.isfinite(out).all()
    out.float().square().mean().backward()
    assert all(p.grad is not None for p in model.parameters() if p.requires_grad)


def test_fused_variants_keep_the_state_dict():
    for name in ("resnet50", "densenet121"):
        with torch.device("meta"):
            a, b = create(name), create(name, fused_bn=True)
        assert list(a.state_dict().keys()) == list(b.state_dict().keys())
    from dear_pytorch_b200.models import bert
    cfg = bert.BertConfig(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128, vocab_size=100)
    with torch.device("meta"):
        a, b = bert.BertForPreTraining(cfg), bert.BertForPreTraining(cfg, fused_ln=True, tc_ffn=True)
    assert list(a.state_dict().keys()) == list(b.state_dict().keys())
    with torch.device("meta"):
        large = bert.BertForPreTraining(bert.BERT_LARGE)
    assert sum(p.numel() for p in large.parameters()) == 336_232_258      # the count bench.py reports for BERT-large


def test_bert_matches_transformers_bert_for_pretraining():
    """Same architecture as the model the reference trains (dear/bert_benchmark.py:72-83): load the weights of a small
    ``transformers.BertForPreTraining`` through the converter and compare both heads, with and without key padding; the
    converter round-trips."""
    transformers = pytest.importorskip("transformers")
    from dear_pytorch_b200.models import bert as B
    cfg = dict(vocab_size=90, hidden_size=32, num_hidden_layers=3, num_attention_heads=4, intermediate_size=64,
               max_position_embeddings=24)
    torch.manual_seed(0)
    hf = transformers.BertForPreTraining(transformers.BertConfig(**cfg)).eval()
    with torch.no_grad():                                  # the zero-initialised biases / unit LayerNorms prove nothing
        for p in hf.parameters():
            p.add_(0.05 * torch.randn_like(p))
    ours = B.BertForPreTraining(B.BertConfig(**cfg)).eval()
    assert ours.vocab_size == 96                           # padded to a multiple of 8 like the reference does (:77-78)
    ours.load_state_dict(B.from_hf_state_dict(hf.state_dict(), cfg["num_hidden_layers"], ours.vocab_size))
    ids = torch.randint(0, 90, (3, 20))
    types = torch.randint(0, 2, (3, 20))
    mask = torch.ones(3, 20, dtype=torch.long)
    mask[1, 13:] = 0
    mask[2, 5:] = 0
    with torch.no_grad():
        for m in (None, mask):
            ref = hf(input_ids=ids, token_type_ids=types, attention_mask=m)
            scores, nsp = ours(ids, types, m)
            torch.testing.assert_close(scores[..., :90], ref.prediction_logits, rtol=1e-4, atol=1e-4)
            torch.testing.assert_close(nsp, ref.seq_relationship_logits, rtol=1e-4, atol=1e-4)
    back = B.to_hf_state_dict(ours.state_dict(), cfg["num_hidden_layers"], 90)
    missing = hf.load_state_dict(back, strict=False)
    assert not missing.unexpected_keys and all("position_ids" in k for k in missing.missing_keys)
    for k, v in hf.state_dict().items():
        if k in back:
            assert torch.equal(back[k], v), k
