"""Mirror of the reference's metadata-aware handler base
(/root/reference/rumpy/SISR/models/attention_manipulators/__init__.py:11-193): `QModel` turns the per-image metadata
rows the data pipeline delivers into the [N, M, 1, 1] vector the meta-attention blocks consume and passes it to
`net.forward(x, metadata=...)`.  Same constructor arguments, attributes and method names; the contrastive
(MoCo) encoder and the SFT / channel-concatenation variants are outside the native trunk and raise."""
import os

import numpy as np
import torch

from rumpy_b200.shared_framework.models.base_architecture import BaseModel


class QModel(BaseModel):
    def __init__(self, metadata=None, use_moco=None, pre_trained_encoder_weights=None, metadata_bypass_len=None,
                 ignore_degradation_location=False, **kwargs):
        self.style = None
        self.channel_concat = False
        self.no_metadata = False
        self.metadata_keys_used_in_training = None
        self.ignore_degradation_location = ignore_degradation_location
        if use_moco:
            raise NotImplementedError('rumpy_b200 QModel: the contrastive (MoCo) metadata encoder is not part of the '
                                      'native trunk')
        if metadata_bypass_len:
            self.num_metadata = metadata_bypass_len
            self.metadata = None
        elif metadata is not None:
            # vector sizes per metadata key: reference __init__.py:27-52
            self.num_metadata = len(metadata)
            extra = {'contrastive_encoding': 255, 'contrastive_q': 255, 'contrastive_encoding_tsne': 1,
                     'contrastive_q_tsne': 1, 'contrastive_encoding_pca': 10, 'contrastive_q_pca': 7, 'all': 39}
            for key, n in extra.items():
                if key in metadata:
                    self.num_metadata += n
            if 'blur_kernel' in metadata:
                self.num_metadata += 9
            elif 'unmodified_blur_kernel' in metadata or any('unmodified_blur_kernel' in m for m in metadata):
                self.num_metadata += 440
            self.metadata = metadata
            if self.ignore_degradation_location:
                self.metadata = [m[2:] if m[0].isdigit() else m for m in self.metadata]
        else:
            self.metadata = ['qpi']
            self.num_metadata = 1
        super(QModel, self).__init__(**kwargs)
        self.moco_encoding = False

    def _metadata_table(self, n, metadata, keys):
        """The per-image metadata rows as an [n, k] fp32 CPU table restricted to the keys this model was configured
        with (`self.metadata`; 'all' keeps every column).  With a single key the row is the value itself."""
        if metadata is None:
            raise RuntimeError('Metadata needs to be specified for this network to run properly.')
        rows = metadata if torch.is_tensor(metadata) else torch.as_tensor(np.asarray(metadata))
        table = rows.detach().to('cpu', torch.float32).reshape(n, -1)
        if len(keys) > 1:
            if 'all' in self.metadata:
                keep = torch.ones(self.num_metadata, dtype=torch.bool)
            else:
                keep = torch.tensor([key[0] in self.metadata for key in keys], dtype=torch.bool)
            table = table[:, keep]
        return table

    def generate_channels(self, x, metadata, keys):
        """Per-image metadata rows -> the [N, num_metadata, 1, 1] vector the q-layers take (what the reference's
        per-image loop builds, __init__.py:87-108): one broadcast instead of a Python loop over the batch."""
        n = x.size(0)
        vec = self._metadata_table(n, metadata, keys).expand(n, self.num_metadata).clone()[:, :, None, None]
        return self.scale_qpi(vec) if self.style == 'modulate' else vec

    def generate_sft_channels(self, x, metadata, metadata_keys):
        raise NotImplementedError('rumpy_b200 QModel: SFT / SRMD channel-tiled metadata is outside the native trunk')

    def channel_concat_logic(self, x, extra_channels, metadata, metadata_keys):
        """(input batch, metadata vector) for the network call (reference __init__.py:137-165).  Ready-made
        `extra_channels` pass through; the key names seen first are remembered for the checkpoint."""
        if self.channel_concat:
            raise NotImplementedError('rumpy_b200 QModel: metadata concatenated with the input image (SRMD mode)')
        if self.no_metadata:
            return x, None
        if extra_channels is None:
            extra_channels = self.generate_channels(x, metadata, metadata_keys)
        if self.metadata_keys_used_in_training is None and metadata_keys is not None:
            self.metadata_keys_used_in_training = [key[0] for key in metadata_keys]
        return x, extra_channels

    def save_model(self, model_save_name, extract_state_only=True, minimal=False):
        """Reference __init__.py:167-175: the base class only assembles the state (extract_state_only=True by
        default here), the file is ALWAYS written by this override, with the metadata key names when known."""
        super().save_model(model_save_name=model_save_name, extract_state_only=extract_state_only, minimal=minimal)
        if self.metadata_keys_used_in_training:
            self.state['metadata_keys_used_in_training'] = self.metadata_keys_used_in_training
        torch.save(self.state, f=os.path.join(self.model_save_dir, '{}_{}'.format(model_save_name, self.curr_epoch)))

    def _with_metadata(self, base_call, x, y, metadata, metadata_keys, extra_channels, **kwargs):
        data, vec = self.channel_concat_logic(x, extra_channels, metadata, metadata_keys)
        return base_call(data, y, extra_channels=vec, **kwargs)

    def run_train(self, x, y, metadata=None, extra_channels=None, metadata_keys=None, *args, **kwargs):
        return self._with_metadata(super().run_train, x, y, metadata, metadata_keys, extra_channels, **kwargs)

    def run_eval(self, x, y=None, request_loss=False, metadata=None, metadata_keys=None,
                 extra_channels=None, *args, **kwargs):
        return self._with_metadata(super().run_eval, x, y, metadata, metadata_keys, extra_channels,
                                   request_loss=request_loss, **kwargs)

    def run_model(self, x, extra_channels=None, *args, **kwargs):
        return self.net.forward(x, metadata=extra_channels)
