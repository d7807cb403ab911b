;;;; cuda.lisp -- the :cuda einsum backend: src/4einsum-backends/cuda.lisp in a numcl checkout.
;;;;
;;;; SOURCE ONLY, never loaded (no Lisp implementation in the build image); see numcl-cuda-ffi.lisp.
;;;;
;;;; What changes in numcl:
;;;;   (setf numcl.impl::*compiler* :cuda)             ; src/3einsum.lisp:264
;;;; and these three dispatch points.  Element-type inference (2typeinfer.lisp), shape unification
;;;; (shape-resolver), output allocation (%output-generator) and the public API are untouched.

(in-package :numcl.impl)

;;; 1. The plugin hook: einsum-body dispatched on *compiler* (src/3einsum.lisp:266-272).
;;; The method is called at code-generation time and must return a FORM.  The form is spliced into
;;; einsum-lambda (src/3einsum.lisp:476-493), inside SPECIALIZING, where I0.. and O0.. are bound to
;;; the 1-D base vectors (:485-487) and the ?n loop limits are known.  Our form serialises the
;;; einsum-specs struct once (load time) and calls ncl_einsum with host-resident tensors.

(defun transform->expr-nodes (form)
  "Flatten a TRANSFORM form into the post-order node list of ncl_expr.  N-ary + * are folded to
the left, max/min to the right (SBCL's source transform), (- x) is NEG, (/ x) is (/ 1 x)."
  (let ((nodes nil) (n 0))
    (labels ((emit (op a b ival fval) (push (list op a b ival fval) nodes) (1- (incf n)))
             (binop (name) (binary-op name))          ; numcl-cuda-ffi.lisp: ONE table for every dispatch point
             (unop (name) (unary-op name))
             (rec (f)
               (etypecase f
                 (integer (emit :x-const-int 0 0 f 0d0))
                 (single-float (emit :x-const-f32 0 0 0 (float f 1d0)))
                 (double-float (emit :x-const-f64 0 0 0 f))
                 (symbol
                  (let ((name (symbol-name f)))
                    (cond ((char= #\$ (aref name 0)) (emit :x-input (parse-integer name :start 1) 0 0 0d0))
                          ((char= #\@ (aref name 0)) (emit :x-output (parse-integer name :start 1) 0 0 0d0))
                          (t (error "numcl-cuda: free variable ~a in an einsum transform" f)))))
                 (cons
                  (destructuring-bind (head . args) f
                    (case head
                      ((+ *) (if (null args)
                                 (emit :x-const-int 0 0 (if (eq head '+) 0 1) 0d0)
                                 (reduce (lambda (acc x) (emit (binop head) acc (rec x) 0 0d0))
                                         (rest args) :initial-value (rec (first args)))))
                      ((- /) (if (null (rest args))
                                 (if (eq head '-)
                                     (emit :neg (rec (first args)) 0 0 0d0)
                                     (let ((one (emit :x-const-int 0 0 1 0d0)))
                                       (emit :div one (rec (first args)) 0 0d0)))
                                 (reduce (lambda (acc x) (emit (binop head) acc (rec x) 0 0d0))
                                         (rest args) :initial-value (rec (first args)))))
                      ((max min) (let ((vals (mapcar #'rec args)))
                                   (reduce (lambda (x acc) (emit (binop head) x acc 0 0d0))
                                           (butlast vals) :initial-value (car (last vals)) :from-end t)))
                      ;; (floor number &optional (divisor 1)) etc., src/4numeric.lisp:572-577
                      ((floor ceiling round truncate mod rem)
                       (emit (binop head) (rec (first args))
                             (if (rest args) (rec (second args)) (emit :x-const-int 0 0 1 0d0)) 0 0d0))
                      (t (if (rest args)
                             (emit (binop head) (rec (first args)) (rec (second args)) 0 0d0)
                             (emit (unop head) (rec (first args)) 0 0 0d0)))))))))
      (rec form)
      (nreverse nodes))))

(defun serialise-einsum-specs (ev)
  "einsum-specs (src/3einsum.lisp:44-62) -> a foreign ncl_einsum_desc that lives as long as the
compiled einsum function.  Transforms outside the whitelist signal an error here, at compile time:
there is no CPU fallback inside the :cuda backend."
  (ematch ev
    ((einsum-specs iter-specs transforms i-specs o-specs i-options o-options)
     (let ((d (cffi:foreign-alloc '(:struct ncl-einsum-desc))))
       (cffi:with-foreign-slots ((n-in n-out n-iter) d (:struct ncl-einsum-desc))
         (setf n-in (length i-specs) n-out (length o-specs) n-iter (length iter-specs)))
       (flet ((slot (name) (cffi:foreign-slot-pointer d '(:struct ncl-einsum-desc) name)))
         (loop for x in iter-specs for i from 0 do (setf (cffi:mem-aref (slot 'iter) :int32 i) x))
         (loop for (specs rank-slot spec-slot opts nopt-slot opt-slot)
                 in `((,i-specs in-rank in-spec ,i-options in-nopt in-opt)
                      (,o-specs out-rank out-spec ,o-options out-nopt out-opt))
               do (loop for spec in specs for a from 0
                        do (setf (cffi:mem-aref (slot rank-slot) :int32 a) (length spec))
                           (loop for x in spec for k from 0
                                 do (setf (cffi:mem-aref (slot spec-slot) :int32 (+ (* a 8) k)) x))
                           (let ((n 0))
                             (loop for (key . plist) in (nth a opts)
                                   when (integerp key)
                                     do (let ((o (cffi:mem-aptr (slot opt-slot) '(:struct ncl-iter-opt) (+ (* a 8) n))))
                                          (destructuring-bind (&key (start t) (end t) (step 1)) plist
                                            (cffi:with-foreign-slots ((key has-start has-end start end step) o (:struct ncl-iter-opt))
                                              (setf key key
                                                    has-start (if (eq start t) 0 1) start (if (eq start t) 0 start)
                                                    has-end (if (eq end t) 0 1) end (if (eq end t) 0 end)
                                                    step step)))
                                          (incf n)))
                             (setf (cffi:mem-aref (slot nopt-slot) :int32 a) n))))
         (loop for tr in transforms for k from 0
               do (let* ((e (cffi:mem-aptr (slot 'transform) '(:struct ncl-expr) k))
                         (nodes (transform->expr-nodes tr)))
                    (assert (<= (length nodes) +ncl-max-expr+))
                    (setf (cffi:foreign-slot-value e '(:struct ncl-expr) 'n) (length nodes))
                    (loop for (op a b ival fval) in nodes for i from 0
                          do (let ((nd (cffi:mem-aptr (cffi:foreign-slot-pointer e '(:struct ncl-expr) 'node)
                                                      '(:struct ncl-expr-node) i)))
                               (cffi:with-foreign-slots ((op a b ival fval) nd (:struct ncl-expr-node))
                                 (setf op (cffi:foreign-enum-value 'ncl-op op) a a b b ival ival fval fval)))))))
       d))))

(defmethod einsum-body ((*compiler* (eql :cuda)) ev)
  (let* ((i-avars (i-avars (einsum-specs-i-specs ev)))
         (o-avars (o-avars (einsum-specs-o-specs ev)))
         (n-in (length i-avars))
         (n-out (length o-avars)))
    ;; NB: einsum-lambda has already replaced each array by its base vector and DROPPED the
    ;; displacement offset (src/3einsum.lisp:485-487); elem_offset = 0 reproduces that.  The array
    ;; shapes are re-derived from the ?n loop limits exactly as the CL backend's strides are.
    `(let ((desc (load-time-value (serialise-einsum-specs ',ev))))
       (sb-sys:with-pinned-objects (,@i-avars ,@o-avars)
         (cffi:with-foreign-objects ((tin '(:struct ncl-tensor) ,n-in) (tout '(:struct ncl-tensor) ,n-out))
           ,@(loop for v in i-avars for spec in (einsum-specs-i-specs ev) for k from 0
                   collect `(fill-ncl-tensor (cffi:mem-aptr tin '(:struct ncl-tensor) ,k) ,v
                                             (%cuda-shape ',spec ,(if (member -1 spec) `(aref i-broadcast-shapes ,k) nil)
                                                          (list ,@(mapcar #'? (remove -1 spec))))
                                             0 (array-element-type ,v)))
           ,@(loop for v in o-avars for spec in (einsum-specs-o-specs ev) for k from 0
                   collect `(fill-ncl-tensor (cffi:mem-aptr tout '(:struct ncl-tensor) ,k) ,v
                                             (%cuda-shape ',spec ,(if (member -1 spec) `(first o-broadcast-shapes) nil)
                                                          (list ,@(mapcar #'? (remove -1 spec))))
                                             0 (array-element-type ,v)))
           ;; Outputs were created by the wrapper with ZEROS, or supplied by the caller and must be
           ;; accumulated into (src/3einsum.lisp:468): either way NCL_OUT_ACCUMULATE is exact.
           ;; (A *compiler*-aware %output-generator that allocates with EMPTY and passes
           ;; NCL_OUT_FRESH saves the host zero-fill pass; see INTEGRATION.md.)
           (ncl-check (ncl-einsum desc tin ,n-in tout ,n-out +ncl-out-accumulate+)))))))

(defun %cuda-shape (spec block dims)
  "Array shape as the loop nest sees it: ?n per index, the broadcast block where SPEC has -1."
  (loop for s in spec
        append (if (= s -1) (coerce block 'list) (list (pop dims)))))

;;; 2. broadcast (src/4numeric.lisp:201-294) and the n-ary folds (:528-542).  These never consult
;;; *compiler* in the reference; the glue adds the dispatch at the top of BROADCAST and replaces the
;;; REDUCE over BROADCAST in numcl:+ * - / max min by ONE fused call.

(defun %cuda-call-with-arrays (arrays fn)
  "Pin the base vectors of ARRAYS (numcl arrays), build one ncl_tensor per array, call FN on the block."
  (let ((n (length arrays)))
    (cffi:with-foreign-object (ts '(:struct ncl-tensor) n)
      (labels ((rec (rest k)
                 (if (null rest)
                     (funcall fn ts)
                     (multiple-value-bind (base offset) (array-displacement (first rest))
                       (sb-sys:with-pinned-objects (base)
                         (fill-ncl-tensor (cffi:mem-aptr ts '(:struct ncl-tensor) k) base (shape (first rest))
                                          offset (array-element-type (first rest)))
                         (rec (cdr rest) (1+ k)))))))
        (rec arrays 0)))))

(defun cuda-broadcast (function x y &key type (atomic function))
  "BROADCAST on the :cuda backend: same result type, same result shape, same (assert (broadcast-p ...)).
FUNCTION is any symbol the reference passes (src/4numeric.lisp:528-655: + - * / max min expt, mod rem round floor
ceiling truncate, the ten bitwise functions, the six */bit comparisons); ATOMIC is what two bare numbers are
combined with -- the comparisons pass #'= etc. there so that (numcl:< 1 2) is T, not 1 (:625-630)."
  (when (and (numberp x) (numberp y))
    (return-from cuda-broadcast (funcall atomic x y)))          ; :205-206: scalars never reach a kernel
  (let* ((x (if (arrayp x) x (full nil x)))
         (y (if (arrayp y) y (full nil y)))
         (type (or type (infer-type function (array-element-type x) (array-element-type y)))))
    (assert (broadcast-p x y))
    (let ((r (empty (broadcast-result-shape x y) :type type)))
      (%cuda-call-with-arrays (list x y r)
        (lambda (ts)
          (ncl-check (ncl-broadcast (binary-op function)
                                    (cffi:mem-aptr ts '(:struct ncl-tensor) 0)
                                    (cffi:mem-aptr ts '(:struct ncl-tensor) 1)
                                    (cffi:mem-aptr ts '(:struct ncl-tensor) 2)))))
      r)))

(defun cuda-fold (function use-identity args)
  "numcl:+ etc. as one pass: ((ident op a0) op a1) ... with the element type of every intermediate
array decided exactly as the reference's chain of BROADCAST calls decides it."
  (let* ((arrays (mapcar (lambda (a) (if (arrayp a) a (full nil a))) args))
         (ident-type (strict-type-of (if (member function '(* /)) 1 0)))
         (types (mapcar #'array-element-type arrays))
         (rtype (reduce (lambda (acc ty) (array-element-type (empty 0 :type (infer-type function acc ty))))
                        (if use-identity types (rest types))
                        :initial-value (if use-identity ident-type (first types))))
         (rshape (reduce (lambda (s a) (broadcast-result-shape (empty s) a)) arrays :initial-value nil))
         (r (empty rshape :type rtype)))
    (%cuda-call-with-arrays (append arrays (list r))
      (lambda (ts)
        (ncl-check (ncl-fold (binary-op function)
                             (if use-identity 1 0) ts (length arrays)
                             (cffi:mem-aptr ts '(:struct ncl-tensor) (length arrays))))))
    r))

;;; 3. reduce-array (src/5reduce.lisp:42-65): the compiled nest is replaced by ncl_reduce; the result
;;; array is still (full result-shape initial-element :type type), accumulated into.

(defun cuda-reduce-array (fn array &key axes
                                     (type (%reduce-array-result-type array fn))
                                     (initial-element (zero-value type)))
  (let* ((axes (etypecase axes (null (iota (rank array))) (fixnum (list axes)) (list (sort (copy-list axes) #'<))))
         (remaining (sort (set-difference (iota (rank array)) axes) #'<))
         (result (full (mapcar (lambda (d) (array-dimension array d)) remaining) initial-element :type type)))
    (cffi:with-foreign-object (ax :int (max 1 (length axes)))
      (loop for a in axes for i from 0 do (setf (cffi:mem-aref ax :int i) a))
      (%cuda-call-with-arrays (list array result)
        (lambda (ts)
          (ncl-check (ncl-reduce (ecase fn (+ :add) (* :mul) (max :max) (min :min))
                                 (cffi:mem-aptr ts '(:struct ncl-tensor) 0) ax (length axes)
                                 (cffi:mem-aptr ts '(:struct ncl-tensor) 1) +ncl-out-accumulate+)))))
    (ensure-singleton result)))

;;; 4. numcl:mean / variance / standard-deviation (src/5reduce.lisp:128-167): ONE fused launch instead of
;;; sum -> broadcast '/ -> sqrt.  The reference's definition is kept: variance is the mean of SQUARES.  A full
;;; reduction of an integer array is an exact Lisp rational in the reference, so that case stays on the composition.

(defun cuda-statistic (stat array axes)
  "STAT: 1 mean, 2 variance, 3 standard-deviation (NCL_STAT_*).  Result element type as the reference infers it for
(numcl:/ (numcl:sum ...) count): double-float for double-float arrays, single-float otherwise."
  (let* ((axes (etypecase axes (null (iota (rank array))) (fixnum (list axes)) (list (sort (copy-list axes) #'<))))
         (remaining (sort (set-difference (iota (rank array)) axes) #'<))
         (type (if (csubtypep (array-element-type array) 'double-float) 'double-float 'single-float))
         (result (empty (mapcar (lambda (d) (array-dimension array d)) remaining) :type type)))
    (cffi:with-foreign-object (ax :int (max 1 (length axes)))
      (loop for a in axes for i from 0 do (setf (cffi:mem-aref ax :int i) a))
      (%cuda-call-with-arrays (list array result)
        (lambda (ts)
          (ncl-check (ncl-reduce-stat stat (cffi:mem-aptr ts '(:struct ncl-tensor) 0) ax (length axes)
                                      (cffi:mem-aptr ts '(:struct ncl-tensor) 1))))))
    (ensure-singleton result)))

;;; 5. Residency (north_star: "3array.lisp gains pinned-host and device buffers behind numcl's specialized arrays").
;;; The choke point is %make-array (src/1instantiate.lisp:81-96), which creates every base vector.  With the :cuda
;;; backend it allocates the vector in non-moving foreign memory (static-vectors style: the Lisp array header sits in
;;; front of pinned memory from ncl_host_alloc_pinned), registers it, and unregisters it from a finalizer.  From then on
;;; a call uploads the vector only when the Lisp side wrote into it since the last call.

(defvar *resident-vectors* (make-hash-table :test 'eq :weakness :key)
  "base vector -> its address, for the vectors that carry a device mirror")

(defun cuda-register-base-vector (base &key fresh lazy)
  "Called at the end of %make-array when *compiler* is :cuda.  FRESH: the vector holds nothing yet (EMPTY / a result):
no first upload.  LAZY: results written into it stay on the device until CUDA-FETCH (used for intermediates)."
  (let* ((sap (sb-sys:vector-sap base))
         (bytes (* (ceiling (* (length base) (sb-vm::simple-array-widetag->bits-per-elt (sb-kernel:widetag-of base))) 8))))
    (ncl-check (ncl-host-register sap bytes (logior (if fresh 1 0) (if lazy 2 0))))
    (setf (gethash base *resident-vectors*) (sb-sys:sap-int sap))
    (let ((address (sb-sys:sap-int sap)))
      (sb-ext:finalize base (lambda () (ncl-host-unregister (sb-sys:int-sap address))) :dont-save t))
    base))

(defun cuda-touch (array)
  "The Lisp side wrote into ARRAY ((setf aref), replace, fill ...): numcl's own (setf aref) (src/2aref.lisp:226-272) calls
this; user code that writes through cl:aref directly must call it itself."
  (let ((base (array-displacement array)))
    (when (gethash (or base array) *resident-vectors*)
      (sb-sys:with-pinned-objects (base) (ncl-check (ncl-host-dirty (sb-sys:vector-sap (or base array))))))))

(defun cuda-fetch (array)
  "Bring a lazily kept result back before the Lisp side reads ARRAY (numcl:aref, printing, to-simple-array)."
  (let ((base (or (array-displacement array) array)))
    (when (gethash base *resident-vectors*)
      (ncl-check (ncl-host-fetch (sb-sys:vector-sap base))))
    array))
