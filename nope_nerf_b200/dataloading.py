"""Device-resident frame store: the reference keeps every frame as host fp32
(`dataloading/dataset.py:80-81,141`) and ships two whole frames + two depth maps to the GPU every step
(`model/training.py:170-185`).  At millisecond steps that traffic dominates, so all V frames (24.9 MB each at 1080p)
and DPT maps live in HBM once -- 5.2 GB for a 200-frame Tanks scene out of 180 GB -- and `__getitem__` hands out VIEWS
under the reference's keys (`dataloading/dataloading.py:120-139`, `dataset.py:160-200`), already shaped the way
`collate_remove_none` + `DataLoader(batch_size=1)` would deliver them.  `Trainer.train_step` gathers the sampled pixels
straight from the view (zero copies).  With `device='cpu'` the store is page-locked host memory: same views, and the
loss kernel gathers the sampled pixels in place over PCIe."""
import random

import torch


class ResidentDataset(torch.utils.data.Dataset):
    """imgs (V,3,H,W) fp32 in [0,1], dpts (V,h_d,w_d) fp32 DPT depth priors, K (4,4) camera matrix (`dataset.py:101-104`).

    load_ref_img / random_ref follow `DataField.load_ref_img` (`dataset.py:168-188`): the reference frame of view i is
    i + randint(1, min(random_ref, V-i-1)), and i-1 for the last view."""

    def __init__(self, imgs, dpts, K, device="cuda", load_ref_img=False, random_ref=1, pin_host=True):
        imgs = torch.as_tensor(imgs, dtype=torch.float32); dpts = torch.as_tensor(dpts, dtype=torch.float32)
        if imgs.dim() != 4 or imgs.shape[1] != 3 or dpts.dim() != 3 or dpts.shape[0] != imgs.shape[0]:
            raise ValueError("imgs must be (V,3,H,W) and dpts (V,h_d,w_d); got %s and %s" % (tuple(imgs.shape), tuple(dpts.shape)))
        device = torch.device(device)
        if device.type == "cpu":
            imgs, dpts = imgs.contiguous(), dpts.contiguous()
            if pin_host and torch.cuda.is_available():
                imgs, dpts = imgs.pin_memory(), dpts.pin_memory()
        else:
            imgs, dpts = imgs.to(device).contiguous(), dpts.to(device).contiguous()
        self.imgs, self.dpts = imgs, dpts
        self.K = torch.as_tensor(K, dtype=torch.float32).reshape(1, 4, 4).clone()          # stays on the host (Trainer reads its diagonal)
        self.scale_mat = torch.eye(4).reshape(1, 4, 4)
        self.load_ref_img, self.random_ref = bool(load_ref_img), int(random_ref)
        self.n_views = imgs.shape[0]
        if self.load_ref_img and (self.n_views < 2 or self.random_ref < 1):
            raise ValueError("reference frames need at least two views and random_ref >= 1")

    @classmethod
    def from_reference_field(cls, field, device="cuda", pin_host=True):
        """Resident copy of the reference's loaded scene: `field` is the `DataField` that `dataloading.get_dataloader(cfg)` returns as
        `fields['img']` (`dataloading/dataloading.py:13-45`, `dataset.py:18-153`): `.imgs` (V,3,H,W), `.dpt_depth` (V,[1,]h_d,w_d),
        `.K`, `.ref_img` / `.random_ref`.  Items then equal what the reference's `DataLoader(batch_size=1)` collates (same keys, shapes,
        dtypes), minus the host->device copy per step."""
        if getattr(field, "dpt_depth", None) is None:
            raise ValueError("the reference field holds no DPT depth maps (use_DPT=True runs the depth network inside the step: out of scope)")
        imgs = torch.as_tensor(field.imgs, dtype=torch.float32)
        dpts = torch.as_tensor(field.dpt_depth, dtype=torch.float32)
        dpts = dpts.reshape(dpts.shape[0], dpts.shape[-2], dpts.shape[-1])
        ds = cls(imgs, dpts, field.K, device=device, load_ref_img=bool(field.ref_img), random_ref=int(field.random_ref) or 1, pin_host=pin_host)
        return ds

    def __len__(self):
        return self.n_views

    def ref_index(self, idx, rng=random):
        if idx == self.n_views - 1:
            return idx - 1
        return idx + rng.randint(1, min(self.random_ref, self.n_views - idx - 1))

    def __getitem__(self, idx):
        idx = int(idx)
        if not 0 <= idx < self.n_views:
            raise IndexError(idx)
        data = {"img": self.imgs[idx:idx + 1], "img.idx": torch.tensor([idx]), "img.dpt": self.dpts[idx:idx + 1],
                "img.camera_mat": self.K, "img.scale_mat": self.scale_mat}
        if self.load_ref_img:
            j = self.ref_index(idx)
            data["img.ref_imgs"] = self.imgs[j:j + 1]; data["img.ref_dpts"] = self.dpts[j:j + 1]; data["img.ref_idxs"] = torch.tensor([j])
        return data

    def batches(self, shuffle=True, generator=None):
        """one epoch of collated batch-size-1 items (what `for batch in train_loader` yields in train.py:180), no DataLoader
        workers, no copies"""
        order = torch.randperm(self.n_views, generator=generator).tolist() if shuffle else range(self.n_views)
        for i in order:
            yield self[i]
