"""Embedding API of the pretrained ATST-Frame encoder - audiossl/methods/atstframe/embedding.py:19-127
(SURVEY.md section 8f, row f3): ``load_model``, ``get_scene_embedding``, ``get_timestamp_embedding`` with the
reference's signatures, chunking (1001 frames = the positional-embedding capacity), layer count and output layout.

Everything runs on the audio's device with the training kernels: the batched fused log-mel kernel replaces the
MelSpectrogram -> AmplitudeToDB -> MinMax Compose (embedding.py:12-15), each chunk is one pass of the CUDA encoder
collecting the final-norm outputs of the last 12 blocks.  There is no CPU path: audio on the host raises."""
import torch

from ...transforms.mel import LogMelSpectrogram
from .model import FrameATSTLightningModule

N_BLOCKS = 12
CHUNK_LEN = 1001  # 10 seconds, the length of the positional embedding


def load_model(model_path):
    """the teacher encoder of a Lightning checkpoint of FrameATSTLightningModule (the layout the reference's own
    training writes: ``state_dict`` + ``hyper_parameters``), in eval mode, annotated like embedding.py:19-38."""
    s = torch.load(model_path, map_location="cpu", weights_only=False)
    pretrained_model = FrameATSTLightningModule.load_from_checkpoint(model_path)
    pretrained_encoder = pretrained_model.model.teacher.encoder
    pretrained_encoder.hyper_param = s['hyper_parameters']
    pretrained_encoder.sample_rate = 16000
    pretrained_encoder.scene_embedding_size = pretrained_encoder.embed_dim * 2 * N_BLOCKS
    pretrained_encoder.timestamp_embedding_size = pretrained_encoder.embed_dim * N_BLOCKS
    pretrained_encoder.eval()
    win = s['hyper_parameters'].get("win_length", 1024) if isinstance(s['hyper_parameters'], dict) else 1024
    pretrained_encoder.transform = LogMelSpectrogram(win_length=win)
    return pretrained_encoder


def _mel_chunks(audio, model):
    if audio.dim() == 2:
        audio = audio.unsqueeze(1)
    else:
        assert audio.dim() == 3
    if not audio.is_cuda:
        raise RuntimeError("audiossl_b200 has no CPU path: move the audio to a B200 (cuda) device")
    model.to(audio.device)
    mel = model.transform(audio)  # [B,1,64,T], top_db clamp per clip
    total_len = mel.shape[-1]
    num_chunks = total_len // CHUNK_LEN + 1
    for i in range(num_chunks):
        start, end = i * CHUNK_LEN, min((i + 1) * CHUNK_LEN, total_len)
        if end > start:
            chunk = mel[:, :, :, start:end]
            yield chunk, torch.full((mel.shape[0],), end - start, dtype=torch.int64, device=audio.device)


@torch.no_grad()
def get_scene_embedding(audio, model):
    """audio [1,N] or [B,1,N] -> [B, N_BLOCKS*emb_size]: per chunk the length-masked frame mean of each of the last
    12 blocks, averaged over the chunks (embedding.py:41-82)."""
    output = [model.get_intermediate_layers(mel_chunk, len_chunk, n=N_BLOCKS) for mel_chunk, len_chunk in
              _mel_chunks(audio, model)]
    return torch.mean(torch.stack(output, dim=0), dim=0)


@torch.no_grad()
def get_timestamp_embedding(audio, model):
    """audio [1,N] or [B,1,N] -> ([B, T, N_BLOCKS*emb_size], timestamps [B, T] in ms, one frame per 40 ms)
    (embedding.py:85-127)."""
    B = audio.shape[0]
    output = [model.get_intermediate_layers(mel_chunk, len_chunk, n=N_BLOCKS, scene=False) for mel_chunk, len_chunk in
              _mel_chunks(audio, model)]
    output = torch.cat(output, dim=1)
    length = output.shape[1]
    timestamps = (torch.arange(length) * 40).float().unsqueeze(0).expand(B, -1)
    return output, timestamps
