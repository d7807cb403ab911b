;;;; numcl-cuda-ffi.lisp -- CFFI bindings of libnumcl_cuda.so (include/numcl_cuda.h).
;;;;
;;;; This file and cuda.lisp are the reference-side half of the drop-in.  They are delivered as
;;;; SOURCE ONLY: no Common Lisp implementation exists in the build image or on the GPU boxes, so
;;;; they have never been loaded.  They are kept thin on purpose; every decision with numerical
;;;; consequences lives behind the C ABI, where it is tested (tests/ drive the same entry points
;;;; through ctypes).
;;;;
;;;; Load order: after src/3einsum.lisp (needs EINSUM-BODY, EINSUM-SPECS, *COMPILER*), before
;;;; src/4numeric.lisp -- i.e. as src/4einsum-backends/cuda-ffi.lisp; asd-generator picks up every
;;;; file of that directory (asd-generator-data.asd:26-27).  Add :cffi to numcl.asd :depends-on.

(in-package :numcl.impl)

(cffi:define-foreign-library libnumcl-cuda
  (:unix (:or "libnumcl_cuda.so" "/usr/local/lib/libnumcl_cuda.so")))

(defun ensure-cuda-backend (&optional (device 0))
  "Load the library and bind this image to one GPU (one process per GPU)."
  (cffi:use-foreign-library libnumcl-cuda)
  (cffi:with-foreign-object (devs :int 1)
    (setf (cffi:mem-aref devs :int 0) device)
    (ncl-check (cffi:foreign-funcall "ncl_init" :int 1 :pointer devs :int))))

;;; ---- constants mirrored from numcl_cuda.h -----------------------------------------------
(defconstant +ncl-max-rank+ 8)
(defconstant +ncl-max-index+ 16)
(defconstant +ncl-max-operands+ 8)
(defconstant +ncl-max-expr+ 48)
(defconstant +ncl-max-opts+ 8)
(defconstant +ncl-out-accumulate+ 0)
(defconstant +ncl-out-fresh+ 1)

(cffi:defcenum ncl-dtype
  :bit :ub2 :ub4 :ub7 :ub8 :ub15 :ub16 :ub31 :ub32 :ub62 :ub63 :ub64
  :sb8 :sb16 :sb32 :fixnum :sb64 :f32 :f64 :c64 :c128)

(cffi:defcenum ncl-op
  (:add 0) :sub :mul :div :max :min :eq :ne :le :ge :lt :gt :logand :logior :logxor
  :logandc1 :logandc2 :logeqv :lognand :lognor :logorc1 :logorc2
  :floor :ceiling :round :truncate :mod :rem :expt
  (:neg 32) :abs :square :sqrt :lognot :identity :exp :log :sin :cos :tan :tanh :signum
  :asin :acos :atan :sinh :cosh :asinh :acosh :atanh :log2 :isqrt :logcount :integer-length
  :conjugate :realpart :imagpart
  (:x-input 64) :x-output :x-const-int :x-const-f32 :x-const-f64)

(defparameter *binary-ops*
  '((+ . :add) (- . :sub) (* . :mul) (/ . :div) (max . :max) (min . :min)
    (=/bit . :eq) (/=/bit . :ne) (<=/bit . :le) (>=/bit . :ge) (</bit . :lt) (>/bit . :gt)
    (logand . :logand) (logior . :logior) (logxor . :logxor) (logandc1 . :logandc1) (logandc2 . :logandc2)
    (logeqv . :logeqv) (lognand . :lognand) (lognor . :lognor) (logorc1 . :logorc1) (logorc2 . :logorc2)
    (floor . :floor) (ceiling . :ceiling) (round . :round) (truncate . :truncate) (mod . :mod) (rem . :rem)
    (expt . :expt))
  "Every function numcl hands to BROADCAST (src/4numeric.lisp:528-655) -> ncl_op.  ONE table for the einsum transform
compiler, cuda-broadcast and cuda-fold, so that the three dispatch points cannot drift apart.")

(defparameter *unary-ops*
  '((abs . :abs) (%square . :square) (sqrt . :sqrt) (lognot . :lognot) (identity . :identity) (exp . :exp)
    (log . :log) (%logr . :log) (sin . :sin) (cos . :cos) (tan . :tan) (tanh . :tanh) (signum . :signum)
    (asin . :asin) (acos . :acos) (atan . :atan) (sinh . :sinh) (cosh . :cosh) (asinh . :asinh) (acosh . :acosh)
    (atanh . :atanh) (%log2 . :log2) (isqrt . :isqrt) (logcount . :logcount) (integer-length . :integer-length)
    (conjugate . :conjugate) (realpart . :realpart) (imagpart . :imagpart))
  "The functions of define-simple-mapper (src/4numeric.lisp:466-519) -> ncl_op.")

(defun binary-op (name)
  (or (cdr (assoc name *binary-ops*))
      (error "numcl-cuda: ~a is not on the CUDA whitelist of binary functions (there is no CPU fallback)" name)))
(defun unary-op (name)
  (or (cdr (assoc name *unary-ops*))
      (error "numcl-cuda: ~a is not on the CUDA whitelist of unary functions (there is no CPU fallback)" name)))

;;; ---- structs ------------------------------------------------------------------------------
(cffi:defcstruct ncl-expr-node
  (op :int32) (a :int32) (b :int32) (pad :int32) (ival :int64) (fval :double))

(cffi:defcstruct ncl-expr
  (n :int32) (pad :int32) (node (:struct ncl-expr-node) :count 48))

(cffi:defcstruct ncl-iter-opt
  (key :int32) (has-start :int32) (has-end :int32) (pad :int32)
  (start :int64) (end :int64) (step :int64))

(cffi:defcstruct ncl-einsum-desc
  (n-in :int32) (n-out :int32) (n-iter :int32)
  (iter :int32 :count 16)
  (in-rank :int32 :count 8)
  (in-spec :int32 :count 64)            ; [8][8]
  (out-rank :int32 :count 8)
  (out-spec :int32 :count 64)
  (in-nopt :int32 :count 8)
  (out-nopt :int32 :count 8)
  (in-opt (:struct ncl-iter-opt) :count 64)   ; [8][8]
  (out-opt (:struct ncl-iter-opt) :count 64)
  (transform (:struct ncl-expr) :count 8))

(cffi:defcstruct ncl-tensor
  (buf :pointer) (host :pointer) (dtype :int32) (rank :int32)
  (dims :int64 :count 8) (elem-offset :int64))

;;; ---- functions --------------------------------------------------------------------------------
(cffi:defcfun "ncl_last_error" :string)
(cffi:defcfun "ncl_sync" :int)
(cffi:defcfun "ncl_alloc" :pointer (bytes :size) (dev-slot :int))
(cffi:defcfun "ncl_free" :int (buf :pointer))
(cffi:defcfun "ncl_h2d" :int (dst :pointer) (dst-off :size) (host :pointer) (bytes :size))
(cffi:defcfun "ncl_d2h" :int (host :pointer) (src :pointer) (src-off :size) (bytes :size))
(cffi:defcfun "ncl_host_alloc_pinned" :pointer (bytes :size))
(cffi:defcfun "ncl_host_free" :int (p :pointer))
(cffi:defcfun "ncl_einsum" :int
  (desc :pointer) (in :pointer) (n-in :int) (out :pointer) (n-out :int) (flags :unsigned-int))
(cffi:defcfun "ncl_broadcast" :int (op ncl-op) (x :pointer) (y :pointer) (r :pointer))
(cffi:defcfun "ncl_fold" :int (op ncl-op) (use-identity :int) (args :pointer) (n :int) (r :pointer))
(cffi:defcfun "ncl_map" :int (fn :pointer) (args :pointer) (n :int) (r :pointer))
(cffi:defcfun "ncl_reduce" :int
  (op ncl-op) (a :pointer) (axes :pointer) (n-axes :int) (r :pointer) (flags :unsigned-int))
(cffi:defcfun "ncl_convert" :int (a :pointer) (r :pointer))
(cffi:defcfun "ncl_reduce_stat" :int (stat :int) (a :pointer) (axes :pointer) (n-axes :int) (r :pointer))
;; residency: a device mirror attached to a base vector in %make-array (src/1instantiate.lisp:81-96)
(cffi:defcfun "ncl_host_register" :int (base :pointer) (bytes :size) (flags :unsigned-int))
(cffi:defcfun "ncl_host_unregister" :int (base :pointer))
(cffi:defcfun "ncl_host_dirty" :int (base :pointer))
(cffi:defcfun "ncl_host_fetch" :int (base :pointer))
;; CUDA graphs: a recorded call sequence
(cffi:defcfun "ncl_capture_begin" :int)
(cffi:defcfun "ncl_capture_end" :int (out :pointer))
(cffi:defcfun "ncl_graph_launch" :int (graph :pointer))
(cffi:defcfun "ncl_graph_free" :int (graph :pointer))
(cffi:defcfun "ncl_trim" :int)

(define-condition numcl-cuda-error (error)
  ((code :initarg :code :reader numcl-cuda-error-code)
   (message :initarg :message :reader numcl-cuda-error-message))
  (:report (lambda (c s)
             (format s "numcl-cuda error ~a: ~a" (numcl-cuda-error-code c) (numcl-cuda-error-message c)))))

(define-condition numcl-cuda-domain-error (numcl-cuda-error arithmetic-error) ()
  (:documentation "NCL_ERR_DOMAIN (-7): an element left the real domain of its function -- where the :common-lisp
backend would have signalled DIVISION-BY-ZERO or stored a complex number.  Reported at the first synchronisation point."))

(declaim (inline ncl-check))
(defun ncl-check (rc)
  "Every ncl_* returns 0 or a negative status; nothing aborts or throws across the ABI.
-2 (NCL_ERR_SHAPE) corresponds to the reference's failed (assert ...) forms, -7 to its arithmetic errors."
  (unless (zerop rc)
    (error (if (= rc -7) 'numcl-cuda-domain-error 'numcl-cuda-error) :code rc :message (ncl-last-error)))
  rc)

;;; ---- element types ------------------------------------------------------------------------------
(defun ncl-dtype-of (element-type)
  "The ncl_dtype of an SBCL specialized array element type (the storage format is SBCL's own:
packed bits, tagged fixnums, ...: nothing is converted on the way to the device)."
  (cond ((csubtypep element-type 'single-float) :f32)
        ((csubtypep element-type 'double-float) :f64)
        ((csubtypep element-type '(complex single-float)) :c64)     ; pairs (real, imaginary): SBCL's own layout
        ((csubtypep element-type '(complex double-float)) :c128)
        ((csubtypep element-type 'bit) :bit)
        ((csubtypep element-type '(unsigned-byte 2)) :ub2)
        ((csubtypep element-type '(unsigned-byte 4)) :ub4)
        ((csubtypep element-type '(unsigned-byte 7)) :ub7)
        ((csubtypep element-type '(unsigned-byte 8)) :ub8)
        ((csubtypep element-type '(unsigned-byte 15)) :ub15)
        ((csubtypep element-type '(unsigned-byte 16)) :ub16)
        ((csubtypep element-type '(unsigned-byte 31)) :ub31)
        ((csubtypep element-type '(unsigned-byte 32)) :ub32)
        ((csubtypep element-type '(unsigned-byte 62)) :ub62)
        ((csubtypep element-type '(unsigned-byte 63)) :ub63)
        ((csubtypep element-type '(unsigned-byte 64)) :ub64)
        ((csubtypep element-type '(signed-byte 8)) :sb8)
        ((csubtypep element-type '(signed-byte 16)) :sb16)
        ((csubtypep element-type '(signed-byte 32)) :sb32)
        ((csubtypep element-type 'fixnum) :fixnum)
        ((csubtypep element-type '(signed-byte 64)) :sb64)
        (t (error "numcl-cuda: no device representation for element type ~a" element-type))))

;;; ---- tensors: a numcl array = base vector + displaced header (1instantiate.lisp:81-96) ------------
(defun fill-ncl-tensor (ptr base shape offset element-type)
  "Describe the numcl array (BASE vector, SHAPE, displacement OFFSET) as a host-resident ncl_tensor.
The caller must hold BASE pinned (sb-sys:with-pinned-objects) for the duration of the call."
  (cffi:with-foreign-slots ((buf host dtype rank elem-offset) ptr (:struct ncl-tensor))
    (setf buf (cffi:null-pointer)
          host (sb-sys:vector-sap base)
          dtype (cffi:foreign-enum-value 'ncl-dtype (ncl-dtype-of element-type))
          rank (length shape)
          elem-offset offset))
  (let ((dims (cffi:foreign-slot-pointer ptr '(:struct ncl-tensor) 'dims)))
    (loop for d in shape for i from 0 do (setf (cffi:mem-aref dims :int64 i) d)))
  ptr)
